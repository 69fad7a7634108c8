// Warp-specialised tcgen05 GEMM core for sm_100a.
//
//   D[m, n] = sum_{term} sum_k A_term[m, k] * B_term[n, k]        (bf16 operands, fp32 accumulate in TMEM)
//
// One CTA computes one 128 x BN tile.  192 threads:
//   warp 0      TMA producer  (cp.async.bulk.tensor into a kStages ring of 128B-swizzled tiles)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns per warp, then the Epi functor
// "Terms" accumulate several operand pairs into one accumulator; this is how the split-bf16
// (hi/lo) input projection gets near-fp32 accuracy from bf16 tensor-core passes.
// Operands are K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows contiguous; used by the
// weight-gradient GEMMs whose reduction runs over time*batch).
#pragma once
#include "sm100.cuh"

namespace svb {

constexpr int kBM = 128;
constexpr int kBK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 192;
constexpr int kMaxTerms = 3;

struct __align__(64) GemmOperands {
  CUtensorMap ta[kMaxTerms];
  CUtensorMap tb[kMaxTerms];
  int za[kMaxTerms];             // slab (3rd tensor-map coordinate) of A per term
  int zb[kMaxTerms];
  int nterms;
  int M, N, K;                   // K = reduction length per term
};

template <int BN, int kStages>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // barriers + alignment slack
};

// Epi must provide:  struct Params;  static __device__ void apply(const Params&, int m, int n0, float (&acc)[32]);
// and optionally a per-tile prologue.  (m, n0) are global row / first column of the 32-wide chunk.
template <int BN, int kStages, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(kGemmThreads) tc_gemm_kernel(const __grid_constant__ GemmOperands ops,
                                                               const typename Epi::Params ep) {
  using S = GemmSmem<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int nkb = (ops.K + kBK - 1) / kBK;
  const int iters = nkb * ops.nterms;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int t = 0; t < ops.nterms; ++t) {
        tma_prefetch_desc(&ops.ta[t]);
        tma_prefetch_desc(&ops.tb[t]);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int term = it / nkb;
        const int k0 = (it - term * nkb) * kBK;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[stage], S::kStageBytes);
        if (A_MN) {
          for (int j = 0; j < kBM / 64; ++j)
            tma_load_3d(sa + j * (kBK * 128), &ops.ta[term], &full_bar[stage], m0 + 64 * j, k0, ops.za[term]);
        } else {
          tma_load_3d(sa, &ops.ta[term], &full_bar[stage], k0, m0, ops.za[term]);
        }
        if (B_MN) {
          for (int j = 0; j < BN / 64; ++j)
            tma_load_3d(sb + j * (kBK * 128), &ops.tb[term], &full_bar[stage], n0 + 64 * j, k0, ops.zb[term]);
        } else {
          tma_load_3d(sb, &ops.tb[term], &full_bar[stage], k0, n0, ops.zb[term]);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // K-major: 16 bf16 = 32 B further along the swizzled row.  MN-major: 16 k-rows = 2 KB further.
          const uint64_t da = A_MN ? umma_desc_mnmajor_sw128(sa + k * 2048, kBK * 128)
                                   : umma_desc_kmajor_sw128(sa + k * 32);
          const uint64_t db = B_MN ? umma_desc_mnmajor_sw128(sb + k * 2048, kBK * 128)
                                   : umma_desc_kmajor_sw128(sb + k * 32);
          umma_bf16_ss(tmem_d, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);               // frees the smem slot when these MMAs retire
        if (it == iters - 1) umma_commit(accum_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane_id();
    typename Epi::Tile tile;
    Epi::prologue(ep, tile, m, n0, m < ops.M);
    mbar_wait(accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float acc[32];
      tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + c * 32, acc);
      tmem_ld_wait();
      if (m < ops.M && n0 + c * 32 < ops.N) Epi::apply(ep, tile, m, n0 + c * 32, acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_d);
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();

// bf16 tensor viewed as [d2][d1][d0] (d0 contiguous); box = {64, box_rows, 1}, 128B swizzle, OOB -> 0.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t box_rows);

template <int BN, int kStages, bool A_MN, bool B_MN, class Epi>
cudaError_t launch_tc_gemm(const GemmOperands& ops, const typename Epi::Params& ep, cudaStream_t stream) {
  using S = GemmSmem<BN, kStages>;
  auto kern = tc_gemm_kernel<BN, kStages, A_MN, B_MN, Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((ops.N + BN - 1) / BN, (ops.M + kBM - 1) / kBM);
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(ops, ep);
  return cudaGetLastError();
}

}  // namespace svb
