// Epilogue functors for tc_gemm_kernel.  Each thread owns one accumulator row (TMEM lane) and receives the
// row in chunks of 32 consecutive columns.
#pragma once
#include "sm100.cuh"

namespace svb {

// C[m, n] = acc + bias[n]  (fp32, row-major, ld = row pitch in elements)
struct EpiStoreF32 {
  struct Params {
    float* C;
    const float* bias;   // may be null
    int64_t ldc;
    int N;
    int unpack_H;        // > 0: row m is a packed gate row (see lstm.cuh); store to row gate*H + unit
  };
  struct Tile {};
  static __device__ __forceinline__ void prologue(const Params&, Tile&, int, int, bool) {}
  static __device__ __forceinline__ void apply(const Params& p, Tile&, int m, int n0, float (&acc)[32]) {
    if (p.unpack_H > 0) m = ((m & 31) >> 3) * p.unpack_H + (m >> 5) * 8 + (m & 7);
    float* dst = p.C + (int64_t)m * p.ldc + n0;
    if (n0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (p.bias) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + n0 + j);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        *reinterpret_cast<float4*>(dst + j) = v;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) dst[j] = acc[j] + (p.bias ? p.bias[n0 + j] : 0.0f);
    }
  }
};

}  // namespace svb
