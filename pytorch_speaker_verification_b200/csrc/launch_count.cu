// Launch counter: how many kernels of THIS library a region of host code launches (bench.py's `gpu_launches`).
//
// CUPTI's callback API is subscribed for the runtime launch entry points (cudaLaunchKernel, cudaLaunchKernelExC,
// cudaLaunchCooperativeKernel and their per-thread-stream variants); a launch counts as ours when the host stub it
// names lives in this shared object (dladdr), so torch's own kernels and NCCL's are counted separately.  libcupti is
// opened with dlopen: without it the calls report SVB_ERR_UNSUPPORTED and the caller says so.  Not used on any
// timed path (the subscription is active only between begin and end).
#include <cupti.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include "../../include/svb200.h"

namespace svb {
void set_error(const char* what, cudaError_t e);
namespace {
typedef CUptiResult (*SubscribeFn)(CUpti_SubscriberHandle*, CUpti_CallbackFunc, void*);
typedef CUptiResult (*UnsubscribeFn)(CUpti_SubscriberHandle);
typedef CUptiResult (*EnableFn)(uint32_t, CUpti_SubscriberHandle, CUpti_CallbackDomain, CUpti_CallbackId);
struct Counter {
  void* lib = nullptr;
  SubscribeFn subscribe = nullptr;
  UnsubscribeFn unsubscribe = nullptr;
  EnableFn enable = nullptr;
  CUpti_SubscriberHandle handle = nullptr;
  bool active = false;
  long long ours = 0, other = 0;
  std::map<std::string, long long> names;     // our kernels by (mangled) name
  std::mutex mu;
  char self[1024] = "";
} g_cnt;

bool is_ours(const void* host_func) {
  Dl_info info;
  if (!host_func || !dladdr(host_func, &info) || !info.dli_fname) return false;
  return strcmp(info.dli_fname, g_cnt.self) == 0;
}
void CUPTIAPI on_api(void*, CUpti_CallbackDomain domain, CUpti_CallbackId cbid, const void* cbdata) {
  if (domain != CUPTI_CB_DOMAIN_RUNTIME_API) return;
  const CUpti_CallbackData* cb = static_cast<const CUpti_CallbackData*>(cbdata);
  if (cb->callbackSite != CUPTI_API_ENTER) return;
  const void* func = nullptr;
  switch (cbid) {
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernel_v7000:
      func = static_cast<const cudaLaunchKernel_v7000_params*>(cb->functionParams)->func; break;
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernel_ptsz_v7000:
      func = static_cast<const cudaLaunchKernel_ptsz_v7000_params*>(cb->functionParams)->func; break;
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchCooperativeKernel_v9000:
      func = static_cast<const cudaLaunchCooperativeKernel_v9000_params*>(cb->functionParams)->func; break;
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchCooperativeKernel_ptsz_v9000:
      func = static_cast<const cudaLaunchCooperativeKernel_ptsz_v9000_params*>(cb->functionParams)->func; break;
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernelExC_v11060:
      func = static_cast<const cudaLaunchKernelExC_v11060_params*>(cb->functionParams)->func; break;
    case CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernelExC_ptsz_v11060:
      func = static_cast<const cudaLaunchKernelExC_ptsz_v11060_params*>(cb->functionParams)->func; break;
    default: return;
  }
  std::lock_guard<std::mutex> lock(g_cnt.mu);
  if (is_ours(func)) {
    g_cnt.ours++;
    g_cnt.names[cb->symbolName ? cb->symbolName : "?"]++;
  } else {
    g_cnt.other++;
  }
}
const CUpti_CallbackId kIds[] = {
    CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernel_v7000,           CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernel_ptsz_v7000,
    CUPTI_RUNTIME_TRACE_CBID_cudaLaunchCooperativeKernel_v9000, CUPTI_RUNTIME_TRACE_CBID_cudaLaunchCooperativeKernel_ptsz_v9000,
    CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernelExC_v11060,       CUPTI_RUNTIME_TRACE_CBID_cudaLaunchKernelExC_ptsz_v11060};
}  // namespace
}  // namespace svb
using namespace svb;

extern "C" int svb_launch_count_begin(void) {
  Counter& c = g_cnt;
  if (c.active) return SVB_OK;
  if (!c.lib) {
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&svb_launch_count_begin), &info) && info.dli_fname)
      snprintf(c.self, sizeof(c.self), "%s", info.dli_fname);
    const char* cands[] = {"libcupti.so.12", "libcupti.so", "/usr/local/cuda/targets/x86_64-linux/lib/libcupti.so.12",
                           "/usr/local/cuda/extras/CUPTI/lib64/libcupti.so.12"};
    for (const char* n : cands)
      if ((c.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!c.lib) { set_error("launch counter: libcupti not found", cudaSuccess); return SVB_ERR_UNSUPPORTED; }
    c.subscribe = reinterpret_cast<SubscribeFn>(dlsym(c.lib, "cuptiSubscribe"));
    c.unsubscribe = reinterpret_cast<UnsubscribeFn>(dlsym(c.lib, "cuptiUnsubscribe"));
    c.enable = reinterpret_cast<EnableFn>(dlsym(c.lib, "cuptiEnableCallback"));
    if (!c.subscribe || !c.unsubscribe || !c.enable) { set_error("launch counter: CUPTI symbols missing", cudaSuccess); return SVB_ERR_UNSUPPORTED; }
  }
  if (c.subscribe(&c.handle, on_api, nullptr) != CUPTI_SUCCESS) {
    set_error("launch counter: cuptiSubscribe failed (another CUPTI client is attached)", cudaSuccess);
    return SVB_ERR_UNSUPPORTED;
  }
  for (CUpti_CallbackId id : kIds) c.enable(1, c.handle, CUPTI_CB_DOMAIN_RUNTIME_API, id);
  c.ours = c.other = 0;
  c.names.clear();
  c.active = true;
  return SVB_OK;
}

extern "C" int svb_launch_count_end(long long* ours, long long* other, char* names, size_t names_bytes) {
  Counter& c = g_cnt;
  if (!c.active) return SVB_ERR_ARG;
  c.unsubscribe(c.handle);
  c.active = false;
  if (ours) *ours = c.ours;
  if (other) *other = c.other;
  if (names && names_bytes) {
    std::string s;
    for (const auto& kv : c.names) s += kv.first + "=" + std::to_string(kv.second) + ";";
    snprintf(names, names_bytes, "%s", s.c_str());
  }
  return SVB_OK;
}
