// Fused gradient clipping + SGD update for sm_100a.
//
// Replaces the eager tail of the reference's training step, train_speech_embedder.py:63-65:
//     torch.nn.utils.clip_grad_norm_(embedder_net.parameters(), 3.0)
//     torch.nn.utils.clip_grad_norm_(ge2e_loss.parameters(), 1.0)
//     optimizer.step()                       # torch.optim.SGD, lr = hp.train.lr, no momentum / weight decay (:33-36)
// which stock torch runs as ~40 launches over the 14 + 2 tensors.  Two launches here, HBM-bound:
//   1  sumsq_kernel   : per-block partial sums of g^2 per clip group            (reads 48.5 MB)
//   2  clip_sgd_kernel: total_norm_g = sqrt(sum of partials) in a fixed order (every block re-reduces the few
//                       hundred partials: deterministic, no float atomics), coef_g = min(1, max_norm_g /
//                       (total_norm_g + 1e-6)) as clip_grad_norm_ does, then p -= lr * (coef_g * g), optionally
//                       storing the clipped gradient like clip_grad_norm_ does in place
//                       (reads 2 x 48.5 MB, writes 48.5 MB [+ 48.5 MB])
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace svb {
void set_error(const char* what, cudaError_t e);

constexpr int kOptMaxTensors = 32;
constexpr int kOptMaxGroups = 4;
constexpr int kOptThreads = 256;
constexpr int kOptMaxBlocks = 148 * 4;

struct OptTable {
  float* p[kOptMaxTensors];
  float* g[kOptMaxTensors];
  long long n[kOptMaxTensors];
  int group[kOptMaxTensors];
  float max_norm[kOptMaxGroups];
  int nt, ng, vec_ok;
  float lr;
  int write_grads;
};

__device__ __forceinline__ double warp_sum_o(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partials[g * gridDim.x + block]
__global__ void __launch_bounds__(kOptThreads) sumsq_kernel(const OptTable t, double* __restrict__ partials) {
  __shared__ double red[kOptMaxGroups][kOptThreads / 32];
  float acc[kOptMaxGroups];
#pragma unroll
  for (int g = 0; g < kOptMaxGroups; ++g) acc[g] = 0.f;
  const size_t tid = (size_t)blockIdx.x * kOptThreads + threadIdx.x, nthr = (size_t)gridDim.x * kOptThreads;
  for (int i = 0; i < t.nt; ++i) {
    float s = 0.f;
    const long long n = t.n[i];
    if (t.vec_ok) {
      const float4* g4 = reinterpret_cast<const float4*>(t.g[i]);
      for (size_t k = tid; k < (size_t)(n >> 2); k += nthr) {
        const float4 v = __ldg(g4 + k);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      for (size_t k = (size_t)(n & ~3LL) + tid; k < (size_t)n; k += nthr) s += t.g[i][k] * t.g[i][k];
    } else {
      for (size_t k = tid; k < (size_t)n; k += nthr) s += t.g[i][k] * t.g[i][k];
    }
#pragma unroll
    for (int g = 0; g < kOptMaxGroups; ++g)
      if (t.group[i] == g) acc[g] += s;
  }
#pragma unroll
  for (int g = 0; g < kOptMaxGroups; ++g) {
    const double v = warp_sum_o((double)acc[g]);   // per-thread fp32 sums of ~80 squares, everything above in fp64
    if ((threadIdx.x & 31) == 0) red[g][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < kOptMaxGroups) {
    double v = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) v += red[threadIdx.x][w];
    if ((int)threadIdx.x < t.ng) partials[threadIdx.x * gridDim.x + blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kOptThreads) clip_sgd_kernel(const OptTable t, const double* __restrict__ partials,
                                                               int nparts, float* __restrict__ norms_out) {
  __shared__ double red[kOptThreads / 32];
  __shared__ float coef_s[kOptMaxGroups];
  for (int g = 0; g < t.ng; ++g) {   // same fixed-order reduction in every block
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += kOptThreads) v += partials[g * nparts + i];
    v = warp_sum_o(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
      const float norm = (float)sqrt(tot);
      float c = t.max_norm[g] / (norm + 1e-6f);      // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6)
      if (!(t.max_norm[g] > 0.f)) c = 1.0f;          // max_norm <= 0: clipping disabled for this group
      coef_s[g] = c < 1.0f ? c : 1.0f;               // torch.clamp(clip_coef, max=1.0)
      if (blockIdx.x == 0 && norms_out) norms_out[g] = norm;
    }
  }
  __syncthreads();
  const size_t tid = (size_t)blockIdx.x * kOptThreads + threadIdx.x, nthr = (size_t)gridDim.x * kOptThreads;
  const float nlr = -t.lr;
  for (int i = 0; i < t.nt; ++i) {
    const float c = coef_s[t.group[i]];
    const long long n = t.n[i];
    float* __restrict__ p = t.p[i];
    float* __restrict__ g = t.g[i];
    size_t k0 = 0;
    if (t.vec_ok) {
      float4* p4 = reinterpret_cast<float4*>(p);
      float4* g4 = reinterpret_cast<float4*>(g);
      for (size_t k = tid; k < (size_t)(n >> 2); k += nthr) {
        float4 gv = g4[k], pv = p4[k];
        gv.x *= c; gv.y *= c; gv.z *= c; gv.w *= c;               // grad.mul_(clip_coef)
        pv.x = fmaf(nlr, gv.x, pv.x); pv.y = fmaf(nlr, gv.y, pv.y);   // p.add_(grad, alpha=-lr)
        pv.z = fmaf(nlr, gv.z, pv.z); pv.w = fmaf(nlr, gv.w, pv.w);
        p4[k] = pv;
        if (t.write_grads) g4[k] = gv;
      }
      k0 = (size_t)(n & ~3LL);
    }
    for (size_t k = k0 + tid; k < (size_t)n; k += nthr) {
      const float gv = g[k] * c;
      p[k] = fmaf(nlr, gv, p[k]);
      if (t.write_grads) g[k] = gv;
    }
  }
}

}  // namespace svb

using namespace svb;

extern "C" int svb_clip_sgd_workspace_bytes(size_t* bytes) {
  if (!bytes) return SVB_ERR_ARG;
  *bytes = (size_t)kOptMaxGroups * kOptMaxBlocks * sizeof(double);
  return SVB_OK;
}

extern "C" int svb_clip_sgd(void* const* params, void* const* grads, const int64_t* numel, const int32_t* group,
                            int n_tensors, const float* max_norm, int n_groups, float lr, int write_clipped_grads,
                            float* norms_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !grads || !numel || !group || !max_norm || !workspace || n_tensors < 1 || n_tensors > kOptMaxTensors ||
      n_groups < 1 || n_groups > kOptMaxGroups || workspace_bytes < (size_t)kOptMaxGroups * kOptMaxBlocks * sizeof(double)) {
    set_error("svb_clip_sgd: bad argument (<= 32 tensors, <= 4 clip groups)", cudaSuccess);
    return SVB_ERR_ARG;
  }
  OptTable t{};
  t.nt = n_tensors; t.ng = n_groups; t.lr = lr; t.write_grads = write_clipped_grads ? 1 : 0; t.vec_ok = 1;
  long long total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (!params[i] || !grads[i] || numel[i] < 0 || group[i] < 0 || group[i] >= n_groups) {
      set_error("svb_clip_sgd: bad tensor entry", cudaSuccess);
      return SVB_ERR_ARG;
    }
    t.p[i] = static_cast<float*>(params[i]); t.g[i] = static_cast<float*>(grads[i]);
    t.n[i] = numel[i]; t.group[i] = group[i];
    if ((reinterpret_cast<uintptr_t>(params[i]) | reinterpret_cast<uintptr_t>(grads[i])) & 15) t.vec_ok = 0;
    total += numel[i];
  }
  for (int g = 0; g < n_groups; ++g) t.max_norm[g] = max_norm[g];
  long long want = (total / 4 + kOptThreads - 1) / kOptThreads;
  int grid = (int)(want < 1 ? 1 : want > kOptMaxBlocks ? kOptMaxBlocks : want);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* partials = static_cast<double*>(workspace);
  sumsq_kernel<<<grid, kOptThreads, 0, s>>>(t, partials);
  clip_sgd_kernel<<<grid, kOptThreads, 0, s>>>(t, partials, grid, norms_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_clip_sgd: launch", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
