// Per-device one-time host state shared by the launchers (namespace svb).
#pragma once
#include <cuda_runtime.h>

namespace svb {

// ----------------------------------------------------------------------------- per-device host state
// cudaFuncSetAttribute and the SM count are properties of (kernel, DEVICE): one-time flags are kept per device so that a
// process that uses several GPUs (hp.device = "cuda:1" without torch.cuda.set_device) configures each of them.
constexpr int kMaxDevices = 64;
inline int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}
inline int device_sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device_index();
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}
// true the first time it is called for (flag word, current device)
inline bool first_use_on_device(unsigned long long& done_mask) {
  const unsigned long long bit = 1ull << current_device_index();
  if (done_mask & bit) return false;
  done_mask |= bit;
  return true;
}

}  // namespace svb
