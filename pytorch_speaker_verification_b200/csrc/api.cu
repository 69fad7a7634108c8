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

__global__ void scale3_kernel(float* a, size_t na, float* b, size_t nb, float* c, size_t nc, const float* g) {
  const float s = *g;
  const size_t n = na + nb + nc;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (i < na) a[i] *= s;
    else if (i < na + nb) b[i - na] *= s;
    else c[i - na - nb] *= s;
  }
}
extern "C" int svb_scale3(float* a, size_t na, float* b, size_t nb, float* c, size_t nc, const float* g, void* stream) {
  const size_t n = na + nb + nc;
  if (!g || n == 0) return SVB_ERR_ARG;
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  scale3_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, na, b, nb, c, nc, g);
  return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}
