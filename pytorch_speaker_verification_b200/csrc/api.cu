// Error bookkeeping for the C ABI.
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

namespace svb {
void set_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
}  // namespace svb

extern "C" int svb_arch(void) { return 100; }
extern "C" const char* svb_last_error(void) { return g_err; }
