// TI-SV EER threshold sweep (train_speech_embedder.py:132-149 of the reference) for sm_100a.
//
// One pass over the similarity matrix sim[N, Mv, N] (HBM-bound: N*Mv*N*4 bytes read once): every element is
// bucketed against the ascending float32 thresholds (the reference compares in float32 with the Python double
// rounded to float32); per-speaker integer counts "above threshold" for the whole row block and for
// the diagonal are exact.  A second tiny kernel reproduces the reference's float32 arithmetic bit for bit:
// sequential fp32 sum over speakers (Python sum()), the three sequential divisions, and the strict
// "diff > |FAR-FRR|" scan from diff = 1.
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace svb {
void set_error(const char* what, cudaError_t e);

constexpr int kMaxThr = 128;

__device__ __forceinline__ int bucket_of(float v, const float* thr, int T) {
  // number of thresholds strictly below v  (v > thr[t]  <=>  t < bucket); NaN -> 0
  int lo = 0, hi = T;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v > thr[mid]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// grid = speakers [i0, i0+n_local); counts[(0|1), i, t]
__global__ void __launch_bounds__(256) eer_count_kernel(const float* __restrict__ sim, int N, int Mv, int Nc,
                                                        const float* __restrict__ thr_g, int T,
                                                        int* __restrict__ cnt_all, int* __restrict__ cnt_diag,
                                                        int speaker0) {
  __shared__ float thr[kMaxThr];
  __shared__ int h_all[kMaxThr + 1];
  __shared__ int h_diag[kMaxThr + 1];
  const int i = blockIdx.x;                 // local speaker index
  const int gi = speaker0 + i;              // global speaker id (its diagonal column)
  for (int t = threadIdx.x; t <= T; t += blockDim.x) {
    if (t < T) thr[t] = thr_g[t];
    h_all[t] = 0;
    h_diag[t] = 0;
  }
  __syncthreads();
  const float* base = sim + (size_t)i * Mv * Nc;
  const float t0 = thr[0];
  const int n = Mv * Nc;
  if ((Nc & 3) == 0 && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    for (int q = threadIdx.x; q < n / 4; q += blockDim.x) {
      const float4 v = __ldg(b4 + q);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (vv[e] > t0) {
          const int idx = q * 4 + e;
          const int bk = bucket_of(vv[e], thr, T);
          atomicAdd(&h_all[bk], 1);
          if (idx % Nc == gi) atomicAdd(&h_diag[bk], 1);
        }
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const float v = base[idx];
      if (v > t0) {
        const int bk = bucket_of(v, thr, T);
        atomicAdd(&h_all[bk], 1);
        if (idx % Nc == gi) atomicAdd(&h_diag[bk], 1);
      }
    }
  }
  __syncthreads();
  // count(sim > thr[t]) = sum_{bk > t} hist[bk]
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int sa = 0, sd = 0;
    for (int bk = t + 1; bk <= T; ++bk) { sa += h_all[bk]; sd += h_diag[bk]; }
    cnt_all[(size_t)i * T + t] = sa;
    cnt_diag[(size_t)i * T + t] = sd;
  }
}

// out: [0]=EER [1]=threshold index (-1: never selected) [2]=FAR [3]=FRR, then FAR[T], FRR[T]
__global__ void eer_finish_kernel(const int* __restrict__ cnt_all, const int* __restrict__ cnt_diag, int N, int Mv,
                                  int T, float* __restrict__ out) {
  __shared__ float far_s[kMaxThr], frr_s[kMaxThr];
  const int t = threadIdx.x;
  if (t < T) {
    float far_sum = 0.f, frr_sum = 0.f;
    const float mv = (float)Mv;
    for (int i = 0; i < N; ++i) {           // Python sum(): left-to-right float32 adds
      const float ca = (float)cnt_all[(size_t)i * T + t];
      const float cd = (float)cnt_diag[(size_t)i * T + t];
      far_sum = __fadd_rn(far_sum, __fsub_rn(ca, cd));
      frr_sum = __fadd_rn(frr_sum, __fsub_rn(mv, cd));
    }
    const float FAR = __fdiv_rn(__fdiv_rn(__fdiv_rn(far_sum, (float)(N - 1.0)), mv), (float)N);
    const float FRR = __fdiv_rn(__fdiv_rn(frr_sum, mv), (float)N);
    far_s[t] = FAR; frr_s[t] = FRR;
    out[4 + t] = FAR; out[4 + T + t] = FRR;
  }
  __syncthreads();
  if (t == 0) {
    float diff = 1.0f, EER = 0.f, eFAR = 0.f, eFRR = 0.f;
    int sel = -1;
    for (int k = 0; k < T; ++k) {
      const float d = fabsf(__fsub_rn(far_s[k], frr_s[k]));
      if (diff > d) {
        diff = d;
        EER = __fdiv_rn(__fadd_rn(far_s[k], frr_s[k]), 2.0f);
        sel = k; eFAR = far_s[k]; eFRR = frr_s[k];
      }
    }
    out[0] = EER; out[1] = (float)sel; out[2] = eFAR; out[3] = eFRR;
  }
}

}  // namespace svb
using namespace svb;

extern "C" int svb_eer_counts(const float* sim, int n_local, int Mv, int Nc, int speaker0, const float* thresholds,
                              int T, int* cnt_all, int* cnt_diag, void* stream) {
  if (!sim || !thresholds || !cnt_all || !cnt_diag || n_local < 1 || Mv < 1 || Nc < 1 || T < 1 || T > kMaxThr) {
    set_error("svb_eer_counts: bad argument", cudaSuccess);
    return SVB_ERR_ARG;
  }
  eer_count_kernel<<<n_local, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sim, n_local, Mv, Nc, thresholds, T,
                                                                                 cnt_all, cnt_diag, speaker0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_eer_counts", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}

extern "C" int svb_eer_finish(const int* cnt_all, const int* cnt_diag, int N, int Mv, int T, float* out,
                              void* stream) {
  if (!cnt_all || !cnt_diag || !out || N < 2 || T < 1 || T > kMaxThr) return SVB_ERR_ARG;
  eer_finish_kernel<<<1, kMaxThr, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cnt_all, cnt_diag, N, Mv, T, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_eer_finish", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
