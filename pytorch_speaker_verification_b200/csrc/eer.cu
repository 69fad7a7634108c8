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


// Branch-free bucket search for T < 64 (thr padded with +inf up to 64 entries): fixed 6 steps, so the searches of
// the many elements a lane holds are independent instruction streams the scheduler can interleave.
__device__ __forceinline__ int bucket_of64(float v, const float* thr) {
  int lo = 0;
#pragma unroll
  for (int step = 32; step >= 1; step >>= 1) {
    const int c = lo + step;
    if (v > thr[c - 1]) lo = c;
  }
  return lo;
}

// ---- single-launch sweep: one CTA per speaker, integer atomics into the global totals, and the last block to finish
// ---- reproduces the reference's float32 arithmetic.
constexpr int kSweepWarps = 4;
constexpr int kBk = 64;           // bucket rows of the lane-private histogram (T + 1 <= 64 on this path)
// scratch (zeroed once by the host; the last block re-zeroes it): [0] master ticket, then 7 replicas of the totals [7][2][T] and 2T words of group tickets: sum over speakers of
// (cnt_all - cnt_diag) and of cnt_diag, accumulated with integer atomics (exact in any order).
// One CTA (4 warps) per speaker: each warp takes a quarter of the speaker's (Mv x N) row block.
__global__ void __launch_bounds__(32 * kSweepWarps) eer_sweep_kernel(const float* __restrict__ sim, int N, int Mv, int Nc,
                                                                     const float* __restrict__ thr_g, int T,
                                                                     int* __restrict__ cnt_all, int* __restrict__ cnt_diag,
                                                                     unsigned long long* __restrict__ scratch,
                                                                     float* __restrict__ out) {
  __shared__ float thr[kBk];
  // lane-private columns (uncontended increments, no bank conflicts); two 16-bit buckets per word so that a CTA
  // needs 16 KB and all N CTAs are resident in one wave (a lane sees < 65536 elements: checked by the host)
  __shared__ int hist[kSweepWarps][kBk / 2][32];
  __shared__ int wsum[kSweepWarps][kBk];
  __shared__ int h_diag[kBk];
  __shared__ int suf_all[kBk], suf_diag[kBk];   // suffix sums of the buckets
  __shared__ float far_s[kMaxThr], frr_s[kMaxThr];
  __shared__ int last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < kBk; t += blockDim.x) {
    thr[t] = t < T ? thr_g[t] : INFINITY;
    h_diag[t] = 0;
  }
  for (int bk = 0; bk < kBk / 2; ++bk) hist[warp][bk][lane] = 0;
  __syncthreads();
  const int i = blockIdx.x;
  const float* base = sim + (size_t)i * Mv * Nc;
  const float t0 = thr[0];
  const int n = Mv * Nc;
  int (*h)[32] = hist[warp];
  if ((Nc & 3) == 0 && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    const int n4 = n >> 2;
    for (int q0 = 0; q0 < n4; q0 += 128 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = q0 + u * 128 + threadIdx.x;
        v[u] = q < n4 ? __ldg(b4 + q) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // bucket 0 (not above the first threshold: almost every different-speaker similarity) enters no count
          const int bk = bucket_of64(vv[e], thr);
          if (bk) atomicAdd(&h[bk >> 1][lane], 1 << ((bk & 1) * 16));
        }
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const int bk = bucket_of64(base[idx], thr);
      if (bk) atomicAdd(&h[bk >> 1][lane], 1 << ((bk & 1) * 16));
    }
  }
  for (int m = threadIdx.x; m < Mv; m += blockDim.x) {      // the diagonal column of this speaker's row block
    const float v = base[(size_t)m * Nc + i];
    if (v > t0) atomicAdd(&h_diag[bucket_of64(v, thr)], 1);
  }
  __syncwarp();
  for (int bk = lane; bk <= T; bk += 32) {                  // fold the 32 lane-private columns of each bucket
    int tot = 0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) tot += (h[bk >> 1][(l + lane) & 31] >> ((bk & 1) * 16)) & 0xffff;
    wsum[warp][bk] = tot;
  }
  __syncthreads();
  // count(sim > thr[t]) = sum_{bk > t} hist[bk]: suffix sums of the 64 buckets by warp shuffles (warp 0: all elements,
  // warp 1: the diagonal) -- the per-threshold loops over the buckets were the longest part of a CTA's chain
  if (warp < 2) {
    int lo = warp == 0 ? wsum[0][lane] + wsum[1][lane] + wsum[2][lane] + wsum[3][lane] : h_diag[lane];
    int hi = warp == 0 ? wsum[0][32 + lane] + wsum[1][32 + lane] + wsum[2][32 + lane] + wsum[3][32 + lane] : h_diag[32 + lane];
    if (lane > T) lo = 0;
    if (32 + lane > T) hi = 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int yl = __shfl_down_sync(0xffffffffu, lo, o), yh = __shfl_down_sync(0xffffffffu, hi, o);
      if (lane + o < 32) { lo += yl; hi += yh; }
    }
    lo += __shfl_sync(0xffffffffu, hi, 0);
    int* suf = warp == 0 ? suf_all : suf_diag;
    suf[lane] = lo; suf[32 + lane] = hi;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int sa = t + 1 < kBk ? suf_all[t + 1] : 0;
    const int sd = t + 1 < kBk ? suf_diag[t + 1] : 0;
    cnt_all[(size_t)i * T + t] = sa;
    cnt_diag[(size_t)i * T + t] = sd;
    unsigned long long* part = scratch + 1 + (size_t)(blockIdx.x % 7) * 2 * T;
    if (sa - sd) atomicAdd(&part[t], (unsigned long long)(sa - sd));
    if (sd) atomicAdd(&part[T + t], (unsigned long long)sd);
  }
  __threadfence();
  __syncthreads();
  // Two-level block ticket: gridDim.x returning atomics on ONE word serialise in L2 (1024 of them were a large part of
  // the kernel's 5.5 us tail); the blocks take a ticket in one of G groups (the words of the unused 8th replica), the
  // last block of a group takes the master ticket.
  if (threadIdx.x == 0) {
    const unsigned G = 2u * T < 32u ? 2u * T : 32u;
    const unsigned g = blockIdx.x % G;
    const unsigned in_group = (gridDim.x - g + G - 1) / G;              // blocks b with b % G == g
    unsigned long long* grp = scratch + 1 + (size_t)7 * 2 * T;
    int l = 0;
    if (atomicAdd(&grp[g], 1ULL) == in_group - 1) {
      __threadfence();
      const unsigned groups = gridDim.x < G ? gridDim.x : G;
      l = (atomicAdd(&scratch[0], 1ULL) == groups - 1) ? 1 : 0;
    }
    last = l;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  bool exact = true;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    unsigned long long fa = 0, di = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      fa += __ldcg(scratch + 1 + (size_t)r * 2 * T + t);
      di += __ldcg(scratch + 1 + (size_t)r * 2 * T + T + t);
    }
    const unsigned long long fr = (unsigned long long)N * Mv - di;
    // the reference adds float32 terms left to right; integer-valued partial sums are exact below 2^24
    if (fa > (1ULL << 24) || fr > (1ULL << 24)) exact = false;
    const float mv = (float)Mv;
    const float FAR = __fdiv_rn(__fdiv_rn(__fdiv_rn((float)fa, (float)(N - 1.0)), mv), (float)N);
    const float FRR = __fdiv_rn(__fdiv_rn((float)fr, mv), (float)N);
    far_s[t] = FAR; frr_s[t] = FRR;
    out[4 + t] = FAR; out[4 + T + t] = FRR;
  }
  const int all_exact = __syncthreads_and(exact ? 1 : 0);
  for (int t = threadIdx.x; t < 1 + 16 * T; t += blockDim.x) scratch[t] = 0;   // self-cleaning: ready for the next call
  // The reference's strict "diff > |FAR - FRR|" scan from diff = 1 selects the FIRST index of the minimum (if it is
  // below 1): a parallel arg-min with ties to the lower index instead of 50 dependent steps.  T < 64: warps 0 and 1.
  __shared__ float wmin[2];
  __shared__ int wsel[2];
  if (warp < 2) {
    const int t = threadIdx.x;
    float d = INFINITY;
    int sel = -1;
    if (t < T) {
      const float dd = fabsf(__fsub_rn(far_s[t], frr_s[t]));
      if (1.0f > dd) { d = dd; sel = t; }               // (NaN is never selected)
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, o);
      const int os = __shfl_xor_sync(0xffffffffu, sel, o);
      if (os >= 0 && (sel < 0 || od < d || (od == d && os < sel))) { d = od; sel = os; }
    }
    if (lane == 0) { wmin[warp] = d; wsel[warp] = sel; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int sel = wsel[0] >= 0 && (wsel[1] < 0 || wmin[0] <= wmin[1]) ? wsel[0] : wsel[1];
    if (sel >= 0) {
      out[0] = __fdiv_rn(__fadd_rn(far_s[sel], frr_s[sel]), 2.0f);
      out[1] = (float)sel; out[2] = far_s[sel]; out[3] = frr_s[sel];
    } else {
      out[0] = 0.f; out[1] = -1.0f; out[2] = 0.f; out[3] = 0.f;
    }
    if (!all_exact) out[1] = -2.0f;                     // caller falls back to the sequential float32 kernel
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

/* One launch: per-speaker counts + totals + the reference's float32 FAR/FRR/EER selection.  scratch: 1 + 16T uint64 zeroed by the caller.  out as svb_eer_finish; out[1] == -2 means a count exceeded 2^24 (float32 partial sums no
 * longer exact): call svb_eer_finish on the per-speaker counts instead. */
extern "C" int svb_eer_sweep(const float* sim, int N, int Mv, const float* thresholds, int T, int* cnt_all,
                             int* cnt_diag, unsigned long long* scratch, float* out, void* stream) {
  if (!sim || !thresholds || !cnt_all || !cnt_diag || !scratch || !out || N < 2 || Mv < 1 || T < 1 || T >= kBk ||
      (long long)Mv * N >= 65536LL * 128) {
    set_error("svb_eer_sweep: bad argument (T must be < 64, Mv*N < 2^23)", cudaSuccess);
    return SVB_ERR_ARG;
  }
  eer_sweep_kernel<<<N, 32 * kSweepWarps, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sim, N, Mv, N, thresholds, T, cnt_all, cnt_diag, scratch, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_eer_sweep", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
