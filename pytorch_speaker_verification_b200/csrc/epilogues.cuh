// Epilogue functors for tc_gemm_kernel.  Each thread owns one accumulator row (TMEM lane) and receives the
// row in chunks of 32 consecutive columns; bulk global traffic goes through TMA-staged shared memory.
#pragma once
#include "sm100.cuh"
#include "tc_gemm.cuh"

namespace svb {

// C[m, n] = acc + bias[n]  (fp32, row-major).  use_tma: the 128 x BN tile is staged in 128B-swizzled shared memory
// and written by TMA (full-line coalesced); otherwise (row permutation / odd pitches) direct stores.
template <int BN>
struct EpiStoreF32 {
  struct __align__(64) Params {
    CUtensorMap tc;      // fp32 [M rows, N cols], box {32, 128}, SWIZZLE_128B (use_tma)
    float* C;
    const float* bias;   // may be null
    int64_t ldc;
    int N;
    int unpack_H;        // > 0: row m is a packed gate row (see lstm.cu); store to row gate*H + unit (direct path)
    int use_tma;
    int64_t z_stride;    // elements between the outputs of consecutive blockIdx.z slices (split-K partials, direct path)
    int tma_z;           // TMA path: the map is [kz][M][N] and slice blockIdx.z stores into slab blockIdx.z
  };
  static constexpr int kInBytes = 0;
  static constexpr int kOutBytes = (BN / 32) * 16384;
  static __device__ __forceinline__ void issue_loads(const Params&, uint8_t*, uint64_t*, int, int) {}
  static __device__ __forceinline__ void apply(const Params& p, const uint8_t*, uint8_t* out, int row, int m, int n0,
                                               int c, float (&acc)[32], bool valid) {
    const int nc = n0 + c * 32;
    if (p.use_tma) {
      uint8_t* o = out + c * 16384;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        if (p.bias && nc + 4 * j + 3 < p.N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 4 * j));
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        *reinterpret_cast<float4*>(o + sw128(row, j)) = v;
      }
      return;
    }
    if (!valid) return;
    if (p.unpack_H > 0) m = ((m & 31) >> 3) * p.unpack_H + (m >> 5) * 8 + (m & 7);
    float* dst = p.C + (int64_t)blockIdx.z * p.z_stride + (int64_t)m * p.ldc + nc;
    for (int j = 0; j < 32; ++j)
      if (nc + j < p.N) dst[j] = acc[j] + (p.bias ? p.bias[nc + j] : 0.0f);
  }
  static __device__ __forceinline__ void issue_stores(const Params& p, const uint8_t* out, int m0, int n0) {
    if (!p.use_tma) return;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c)
      if (n0 + c * 32 < p.N) tma_store_3d(&p.tc, out + c * 16384, n0 + c * 32, m0, p.tma_z ? (int)blockIdx.z : 0);
  }
};

// Fills Params for a row-major fp32 output; chooses the TMA path when the layout allows it.
template <int BN>
int make_store_params(typename EpiStoreF32<BN>::Params* p, float* C, const float* bias, int M, int N, int64_t ldc,
                      int unpack_H) {
  memset(p, 0, sizeof(*p));
  p->C = C; p->bias = bias; p->ldc = ldc; p->N = N; p->unpack_H = unpack_H;
  p->use_tma = (unpack_H == 0 && (ldc % 4) == 0 && (N % 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0)) ? 1 : 0;
  if (p->use_tma) return make_tmap(&p->tc, C, 4, (uint64_t)N, (uint64_t)M, 1, (uint64_t)ldc, (uint64_t)ldc * M, 32, 128, 3);
  return 0;
}

}  // namespace svb
