// Persistent recurrent forward kernel (included by lstm.cu after EpiLstmFwd).
//
// One cooperative launch runs all T frames of one LSTM layer.  CTA (slice, tile) owns 128 packed gate columns
// (= 32 hidden units x 4 gates) of a 128-row batch tile:
//   * its W_hh slice [128 x H] bf16 is loaded ONCE by TMA and stays resident in shared memory (196 KB at H=768)
//     as the K-major B operand of every frame's MMA;
//   * per frame the four worker warps copy the tile's h_{t-1} rows (written by the 24 CTAs of the same batch
//     tile) from L2 through registers into TENSOR MEMORY with tcgen05.st, so the A operand of tcgen05.mma comes
//     from TMEM (TS form) and no shared memory is spent on activations;
//   * the accumulator (128 fp32 columns of TMEM) is read back with tcgen05.ld and the LSTM cell is applied in
//     registers (same code as the per-frame kernel's epilogue);
//   * frames are ordered by one monotonically increasing counter per batch tile (release/acquire at gpu scope):
//     only the CTAs sharing a batch tile wait for each other, there is no grid-wide barrier.
// TMEM map (512 columns): [0, H/2) = h tile, two bf16 per column; [H/2, H/2 + 128) = accumulator.
#pragma once
// (included inside namespace svb)

struct PlstmParams {
  CUtensorMap tw;                 // packed W_hh [4H rows, H], K-major, box {64, 128}
  const __nv_bfloat16* h_hi;      // [T+1, B, H] (read slot t, write slot t+1)
  __nv_bfloat16* h_hi_w;
  __nv_bfloat16* h_lo;            // [T+1, B, H] or null
  const float* gin;               // [T, B, 4H]
  float* c;                       // [cslots, B, H]
  __nv_bfloat16* gates;           // [T+1, B, 4H] or null
  float* h_last;                  // [B, H] or null (written at t = T-1)
  unsigned* counters;             // one per batch tile of this launch, zeroed by the host
  int B, T, training;
  int row0, rows;                 // batch rows [row0, row0 + rows) handled by this launch
};

constexpr int kPlstmThreads = 192;

template <int H>
__global__ void __launch_bounds__(kPlstmThreads, 1) plstm_fwd_kernel(const __grid_constant__ PlstmParams p) {
  constexpr int NKB = H / 64;             // 64-wide K blocks
  constexpr int HALF = 4;                 // K blocks per load phase (4 x 32 registers per thread in flight)
  constexpr int NPH = NKB / HALF;
  constexpr int A_COLS = H / 2;
  constexpr uint32_t D_COL = A_COLS;
  static_assert(NKB % HALF == 0 && NPH <= 3 && A_COLS + 128 <= 512, "unsupported hidden size for the persistent kernel");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;                                             // NKB x [128 rows x 64] swizzled
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + NKB * 16384);
  uint64_t* a_full = w_full + 1;                                   // [NPH <= 3]
  uint64_t* acc_full = w_full + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_full + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128;                                  // first packed gate column of the slice
  const int m0 = p.row0 + blockIdx.y * 128;
  const int row_end = p.row0 + p.rows;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < NPH; ++i) mbar_init(&a_full[i], 128);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&p.tw);
      mbar_expect_tx(w_full, NKB * 16384);
      for (int kb = 0; kb < NKB; ++kb) tma_load_3d(wsm + kb * 16384, &p.tw, w_full, kb * 64, n0, 0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    mbar_wait(w_full, 0);
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
      for (int ph = 0; ph < NPH; ++ph) {
        mbar_wait(&a_full[ph], t & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < HALF; ++kk) {
            const int kb = ph * HALF + kk;
            const uint32_t sb = smem_u32(wsm + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem + D_COL, tmem + kb * 32 + k * 8, umma_desc_kmajor_sw128(sb + k * 32), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          }
          if (ph == NPH - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: h -> TMEM, cell epilogue
    const int q = warp & 3;
    const int r = m0 + q * 32 + lane;
    const bool valid = r < row_end;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    unsigned* ctr = p.counters + blockIdx.y;
    const unsigned nslices = gridDim.x;
    const size_t BH = (size_t)p.B * H;
    for (int t = 0; t < p.T; ++t) {
      if (t > 0) {
        if (lane == 0) wait_counter_ge(ctr, nslices * (unsigned)t);
        __syncwarp();
      }
      const uint4* src = reinterpret_cast<const uint4*>(p.h_hi + (size_t)t * BH + (size_t)r * H);
#pragma unroll 1
      for (int ph = 0; ph < NPH; ++ph) {
        uint32_t v[HALF][32];
#pragma unroll
        for (int kk = 0; kk < HALF; ++kk) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 x = make_uint4(0u, 0u, 0u, 0u);
            if (valid) x = __ldcg(src + (ph * HALF + kk) * 8 + j);
            v[kk][4 * j + 0] = x.x; v[kk][4 * j + 1] = x.y; v[kk][4 * j + 2] = x.z; v[kk][4 * j + 3] = x.w;
          }
        }
#pragma unroll
        for (int kk = 0; kk < HALF; ++kk) tmem_st32(tmem + lane_base + (ph * HALF + kk) * 32, v[kk]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_full[ph]);
      }
      // operands of the cell update, requested before the accumulator is complete
      CellDirect::Params ep;
      const int cs_prev = p.training ? t : (t & 1), cs_out = p.training ? t + 1 : ((t + 1) & 1);
      ep.gin = p.gin + (size_t)t * p.B * 4 * H;
      ep.c_prev = p.c + (size_t)cs_prev * BH;
      ep.c_out = p.c + (size_t)cs_out * BH;
      ep.h_hi = p.h_hi_w + (size_t)(t + 1) * BH;
      ep.h_lo = p.h_lo ? p.h_lo + (size_t)(t + 1) * BH : nullptr;
      ep.gates = p.gates ? p.gates + (size_t)t * p.B * 4 * H : nullptr;
      ep.h_f32 = (p.h_last && t == p.T - 1) ? p.h_last : nullptr;
      ep.H = H;
      float4 g[4][8], cpv[8];
      const float4* g4 = reinterpret_cast<const float4*>(ep.gin + (size_t)r * 4 * H + n0);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[0][j] = __ldg(g4 + j); g[1][j] = __ldg(g4 + 8 + j); }
        const float4* c4 = reinterpret_cast<const float4*>(ep.c_prev + (size_t)r * H + (n0 >> 2));
#pragma unroll
        for (int j = 0; j < 8; ++j) cpv[j] = __ldcg(c4 + j);
      }
      mbar_wait(acc_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float acc[32];
        tmem_ld32(tmem + lane_base + D_COL + c * 32, acc);
        if (valid && c + 2 < 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[(c + 2) & 3][j] = __ldg(g4 + (c + 2) * 8 + j);
        }
        tmem_ld_wait();
        if (valid) {
          float pre[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pre[4 * j + 0] = acc[4 * j + 0] + g[c][j].x; pre[4 * j + 1] = acc[4 * j + 1] + g[c][j].y;
            pre[4 * j + 2] = acc[4 * j + 2] + g[c][j].z; pre[4 * j + 3] = acc[4 * j + 3] + g[c][j].w;
          }
          const float cp[8] = {cpv[2 * c].x, cpv[2 * c].y, cpv[2 * c].z, cpv[2 * c].w,
                               cpv[2 * c + 1].x, cpv[2 * c + 1].y, cpv[2 * c + 1].z, cpv[2 * c + 1].w};
          CellDirect::cell(ep, r, n0 + c * 32, pre, cp);
        }
      }
      tc_fence_before();
      __threadfence();                                     // h_t visible at gpu scope before the counter moves
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the four worker warps
      if (threadIdx.x == 64) atomicAdd(ctr, 1u);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int H>
static int launch_plstm_fwd(PlstmParams& p, int n_tiles, cudaStream_t s) {
  constexpr int smem = (H / 64) * 16384 + 1024 + 256;
  auto kern = plstm_fwd_kernel<H>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("plstm: cudaFuncSetAttribute", e); return SVB_ERR_CUDA; }
    configured = true;
  }
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)kern, dim3(H / 32, n_tiles), dim3(kPlstmThreads), args, smem, s);
  if (e != cudaSuccess) { set_error("plstm: cooperative launch", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}

