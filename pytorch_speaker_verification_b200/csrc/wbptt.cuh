// Persistent wavefront BPTT kernel: the time-reversed counterpart of wlstm.cuh (included by lstm.cu, namespace svb).
//
// Replaces, for the whole stack at once, the per-frame BPTT launches and the batched dX GEMMs behind
// train_speech_embedder.py:62 (autograd through nn.LSTM):
//     dh^l_t = dG^l_{t+1} W_hh^l + dG^{l+1}_t W_ih^{l+1}         (top layer: + dL/dh_last at t = T-1)
//     dG^l_t = gate backward(dh^l_t, gates_t, c_t, c_{t-1}, running dL/dc)
// The reductions run over K = 4H = 3072 gate columns, the outputs over H = 768 units, so a [128 units x 768 K]
// weight slice fills the 192 KB of tensor memory that a CTA can devote to the stationary A operand, and a product
// needs (H/128) x 4 = 24 CTAs: clusters of 4 CTAs share one 128-unit tile and split K four ways.
//
//   R(l, u, s)  recurrent product of layer l, unit tile u, K slice s: streams dG^l_{t+1}[64-row tile, K slice] (bf16,
//               TMA -> smem ring) against W_hh^T[tile, slice] in TMEM; the accumulator D[unit (lane), row (column)]
//               is a PARTIAL sum.  CTA s of the cluster owns units [32 s, 32 s + 32) of the tile: every CTA pushes
//               the other three quarters of its partial into the owners' shared memory (st.shared::cluster,
//               lane-contiguous 16-byte stores), owners add the three they receive, the dX tile from X(l+1, u, s)
//               and run the gate backward for their 32 units: dG_t overwrites the gate stash in place (TMA store).
//   X(l, u, s)  input-gradient product of layer l >= 1 (dX^{l-1}_t = dG^l_t W_ih^l), same cluster scheme; the owner
//               writes its reduced [32 units x 64 rows] fp32 tile to a small L2-resident ring in fragment order for
//               R(l-1, u, s), which owns the same units.
//
// Ordering by release/acquire counters as in wlstm.cuh:
//   dcnt[l][j]      += 1 by every R(l, ., .) after its dG stores of (t, j) completed: dG^l_t[tile j] complete at
//                      NS * (T - t)
//   xcnt[l][ns][j]   = T - t by X(l, u, s) (ns = 4 u + s) once dX of (t, j) is written
// 30 clusters of 4 = 120 CTAs, all co-resident (cooperative launch).
#pragma once
#include <stdlib.h>

constexpr int kWbTile = 64;
constexpr int kWbKb = 12;              // 64-wide K blocks per CTA (K slice of 768)
constexpr int kWbKbPerStage = 3;       // 12 MMAs per barrier wait
constexpr int kWbStages = 3;           // ring: 3 x 24 KB (a tile is 4 stages; 5 x 16 KB measured 5 % slower)
constexpr int kWbStageBytes = kWbKbPerStage * kWbTile * 128;
constexpr int kWbXRing = 3;
#ifdef SVB_WB_DEPS
constexpr int kWbDeps = SVB_WB_DEPS;   // (experiment: make ALT=1 ALTFLAGS=-DSVB_WB_DEPS=16)
#else
constexpr int kWbDeps = 8;             // dependency slots: the poller runs up to 8 tiles ahead of the slowest waiter
#endif
// Warp roles.  Every single-thread role has a warp of its own: two roles in one warp run time-sliced, and a lane that
// sleeps in mbarrier.try_wait holds the other one up (poller and input loader shared a warp at first: the TMA producer
// then waited 1200 cycles per tile for dependencies that had been satisfied long before).  The single-thread roles sit
// in the highest warp ids (the sub-partition arbiter prefers the highest eligible warp).
constexpr int kWbMathWarps = 16;       // warps 0-15: split-K reduction + gate backward, thread = (row, 4 units)
constexpr int kWbMathThreads = 32 * kWbMathWarps;
constexpr int kWbWarpXch = kWbMathWarps;   // 4 exchange warps (TMEM lane quarter = warp & 3): accumulator -> staged quarters
constexpr int kWbWarpLoad = kWbWarpXch + 4, kWbWarpPoll = kWbWarpXch + 5, kWbWarpStore = kWbWarpXch + 6,
              kWbWarpTma = kWbWarpXch + 7, kWbWarpMma = kWbWarpXch + 8;
constexpr int kWbThreads = 32 * (kWbWarpMma + 1);     // 25 warps = 800 threads, 72 registers
static_assert(kWbWarpXch % 4 == 0, "exchange warp w must own TMEM lane quarter w & 3");
constexpr int kWbAccCol = 384;
// epilogue tiles.  Three in-place buffers (gates -> dG 2 x 8 KB boxes, running dL/dc): busy from their TMA load through
// the gate math until their TMA store has drained (load latency + math + store ~ 5000 cycles; two buffers bounded the
// tile period).  Two read-only buffers for c_{t-1}.  c_t itself is NOT loaded: tanh(c_t) is recomputed from
// c_t = f c_{t-1} + i g with the stashed activations (no measurable effect on the gradients,
// scripts/precision_study_bwd.py "c_t recomputed"; one 8 KB TMA load per tile and 16 KB of shared memory less).
constexpr int kWbIo = 3;
constexpr int kWbOffG = 0, kWbOffDc = 16384, kWbIoBytes = 24576;       // per in-place buffer
constexpr int kWbCpBytes = 8192;                                       // per c_{t-1} buffer
constexpr int kWbStgBytes = kWbIo * kWbIoBytes + 2 * kWbCpBytes;       // 88 KB
// Split-K exchange.  A partial quarter is staged as fp16 [64 rows][32 units] (4 KB; fp16 = 2^-11 halves the DSMEM
// traffic, the partial sums are O(|dh|) under the gradient scale) and leaves through ONE bulk copy per destination
// (cp.async.bulk.shared::cluster, complete_tx on the receiver's mbarrier): 730 cycles per 3 x 4 KB round against
// 1480 for per-thread st.shared::cluster stores + release arrives (scripts/ubench/dsmem_push.cu), and no warp waits
// for a remote round trip.  The own quarter stays fp32 [64 rows][32 units].
constexpr int kWbQuarterBytes = kWbTile * 32 * 2;
constexpr int kWbRecvBytes = 3 * kWbQuarterBytes;
constexpr int kWbSendBytes = 3 * kWbQuarterBytes;
constexpr int kWbOwnBytes = kWbTile * 32 * 4;
constexpr int kWbSmem = kWbStages * kWbStageBytes + kWbStgBytes + 2 * kWbRecvBytes + 2 * kWbSendBytes + 2 * kWbOwnBytes + 1024 + 1024;
static_assert(kWbSmem <= 227 * 1024, "shared memory budget of the BPTT kernel");

struct __align__(64) WbLayer {
  CUtensorMap t_dg;        // bf16 gates / dG [T+1][B][4H], box {64, 64} SW128: B operand, gate tile load + store
  CUtensorMap t_c;         // fp32 c [T+1][B][H], box {32, 64} SW128
  CUtensorMap t_dc;        // fp32 running dL/dc [1][B][H], box {32, 64} SW128, load + store
  const __nv_bfloat16* whhT;   // [H][4H]
  const __nv_bfloat16* wihT;   // [H][4H] (layers >= 1)
  float* xring;            // dX produced by X(l): [kWbXRing][nt][H/32][64 rows x 32 units] fp32 (layers >= 1)
  float* gbias_ih;         // bias gradients of the layer (reference row order), written at the end of the kernel
  float* gbias_hh;         // b_ih and b_hh enter the pre-activation as a sum: identical gradients
};
struct __align__(64) WbParams {
  WbLayer layer[3];
  unsigned* dcnt;          // [L][nt]
  unsigned* xcnt;          // [L][H/32][nt]
  const float* dh_last;    // [B][H] dL/dh of the top layer's last frame (times the gradient scale, lstm.cu)
  const float* inv_scale;  // device scalar: 1 / gradient scale, applied to the bias gradients on the way out
  long long* trace;
  int ablate;              // debug (SVB_WB_ACCOUNT builds only): parts switched off for timing experiments, results are garbage
  int trace_u, trace_s;    // debug: unit tile (cluster) and cluster rank of the traced CTAs (env SVB_TRACE_U / SVB_TRACE_S, default 0)
  int trace_l;             // debug: layer whose R(l, 0, 0) / X(l, 0, 0) CTAs are traced (env SVB_TRACE_LAYER, default 1)
  int B, T, L, H, nt;
};

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta) {
  const uint32_t addr = map_to_cta(smem_u32(local_bar), cta);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// relaxed variant: signals "I have consumed your data" (the loads it orders have already returned their values; a
// release here costs the signalling warp a cluster-scope fence per tile)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* local_bar, uint32_t cta) {
  const uint32_t addr = map_to_cta(smem_u32(local_bar), cta);
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(SVB_TRYWAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("svb: cluster mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// shared memory of this CTA -> shared memory of a peer CTA; the bytes are credited to the PEER's mbarrier
__device__ __forceinline__ void dsmem_bulk_push(uint32_t dst_cluster_addr, uint32_t src_addr, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src_addr), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}
// Warp-wide wait; experiment: ONE polling lane + __syncwarp
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#ifdef SVB_WB_ONE_LANE_WAIT   // measured: 4.3-5.0 ms against 4.05 ms with every lane polling (the elected lane + __syncwarp adds latency)
  if (lane_id() == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void mbar_wait_cluster_warp(uint64_t* bar, uint32_t parity) {
#ifdef SVB_WB_ONE_LANE_WAIT   // measured: 4.3-5.0 ms against 4.05 ms with every lane polling (the elected lane + __syncwarp adds latency)
  if (lane_id() == 0) mbar_wait_cluster(bar, parity);
  __syncwarp();
#else
  mbar_wait_cluster(bar, parity);
#endif
}
// stores with the offset as an immediate (the staging loops are fully unrolled: one STS per element, no address adds)
template <int kOff> __device__ __forceinline__ void sts_f32_imm(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(addr), "f"(v), "n"(kOff) : "memory");
}
template <int kOff> __device__ __forceinline__ void sts_u16_imm(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0+%2], %1;" ::"r"(addr), "h"(v), "n"(kOff) : "memory");
}
template <int I> struct WbStage {   // element I, I + 1 of a 32-row half: rows I and I + 1 of the lane's unit
  static __device__ __forceinline__ void own(uint32_t d, const float (&v)[32]) {
    sts_f32_imm<I * 128>(d, v[I]);
    sts_f32_imm<(I + 1) * 128>(d, v[I + 1]);
    if constexpr (I + 2 < 32) WbStage<I + 2>::own(d, v);
  }
  static __device__ __forceinline__ void send(uint32_t d, const float (&v)[32]) {
    const uint32_t pk = pack_f16x2(v[I], v[I + 1]);
    sts_u16_imm<I * 64>(d, (uint16_t)(pk & 0xffffu));
    sts_u16_imm<(I + 1) * 64>(d, (uint16_t)(pk >> 16));
    if constexpr (I + 2 < 32) WbStage<I + 2>::send(d, v);
  }
};
__device__ __forceinline__ void math_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * 16) : "memory"); }

// wait accounting (debug; make ALT=1 ALTFLAGS=-DSVB_WB_ACCOUNT, scripts/account_wbptt.py): cycles every role thread spends
// in each of its waits, summed over the kernel, per CTA: trace[2 nt 16 + 256 + 32 cta + slot]
#ifdef SVB_WB_ACCOUNT
#define WB_ACC(k, ...) do { const long long a0_ = clock64(); __VA_ARGS__; wacc[k] += clock64() - a0_; } while (0)
#define WB_ACC_OUT(slot0, n) do { if (p.trace) for (int k_ = 0; k_ < (n); ++k_) p.trace[2 * nt * 16 + 256 + 32 * blockIdx.x + (slot0) + k_] = wacc[k_]; } while (0)
// ablation mask (scripts/ablate_wbptt.py): 1 MMAs, 2 gate math, 4 operand loads, 8 dependencies, 16 exchange staging
// stores, 32 epilogue input loads, 64 DSMEM push, 128 dG / dc stores, 256 the math warps' proxy fence
#define WB_ABL(bit) ((p.ablate & (bit)) != 0)
#else
#define WB_ABL(bit) false
#define WB_ACC(k, ...) do { __VA_ARGS__; } while (0)
#define WB_ACC_OUT(slot0, n) do { } while (0)
#endif

template <int H>
__global__ void __launch_bounds__(kWbThreads, 1) wbptt_kernel(const __grid_constant__ WbParams p) {
  constexpr int NS = H / 32;                 // 32-unit slices per layer (= R CTAs per layer)
  constexpr int NU = H / 128;                // 128-unit tiles (= clusters per product)
  static_assert(H % 128 == 0 && 4 * H == 4 * kWbKb * 64, "the K slices must be 768 wide");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* stg = ring + kWbStages * kWbStageBytes;
  uint8_t* cst = stg + kWbIo * kWbIoBytes;     // the two c_{t-1} buffers
  uint8_t* recv = stg + kWbStgBytes;
  uint8_t* send = recv + 2 * kWbRecvBytes;
  uint8_t* own = send + 2 * kWbSendBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(own + 2 * kWbOwnBytes);
  uint64_t* empty = full + kWbStages;
  uint64_t* acc_full = empty + kWbStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2] 4 exchange warps
  uint64_t* in_full = acc_empty + 2;           // [3] epilogue input tiles landed (tx)
  uint64_t* stg_full = in_full + 3;            // [3] 16 math warps: outputs staged (and c tiles read)
  uint64_t* stg_free = stg_full + 3;           // [3] store thread: in-place buffer reusable
  uint64_t* recv_full = stg_free + 3;          // [2] arming arrive + owner warp + 12 KB of transactions: all four quarters are here
  uint64_t* peer_free = recv_full + 2;         // [2] 3 remote arrives: my three receivers consumed (and re-armed) use k-1
  uint64_t* send_ready = peer_free + 2;        // [2] the owner exchange warp: own quarter staged
  uint64_t* consumed = send_ready + 2;         // [2] math thread 0: recv + own of this use are in registers
  uint64_t* dep_ready = consumed + 2;          // [kWbDeps]
  uint64_t* dep_free = dep_ready + kWbDeps;    // [kWbDeps]
  uint64_t* x_done = dep_free + kWbDeps;       // [2] X: 16 math warps wrote the dX tile
  uint64_t* x_taken = x_done + 2;              // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(x_taken + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = blockIdx.x >> 2;
  const int s = (int)cluster_ctarank();        // K slice and owned unit quarter
  const bool is_R = cl < p.L * NU;
  const int l = is_R ? cl / NU : 1 + (cl - p.L * NU) / NU;
  const int u = is_R ? cl % NU : (cl - p.L * NU) % NU;
  const int ns = 4 * u + s;                    // 32-unit slice index within the layer
  const WbLayer& ly = p.layer[l];
  const int nt = p.nt, T = p.T;
  const int total = T * nt;                   // (< 2^31: checked by the launcher; 32-bit tile counters keep the per-tile bookkeeping short)
  const bool has_x = is_R && l + 1 < p.L;      // dX from the layer above arrives through its ring
  // debug trace: R(l, 0, 0) rows [0, nt), X(l, 0, 0) rows [nt, 2 nt): 16 clock64 stamps per tile of frame T/2
  long long* const trace_cta = (p.trace && l == p.trace_l && u == p.trace_u && s == p.trace_s) ? p.trace + (is_R ? 0 : (size_t)nt * 16) : nullptr;
  const long long t_cta0 = clock64();
#ifdef SVB_WB_ACCOUNT
  long long wacc[6] = {0, 0, 0, 0, 0, 0};
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWbStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
      mbar_init(&recv_full[b], 2);                // the arming arrive (+ 12 KB of transactions) and the owner exchange warp
      mbar_init(&peer_free[b], 3);
      mbar_init(&send_ready[b], 1);
      mbar_init(&consumed[b], 1);
      mbar_init(&x_done[b], kWbMathWarps);
      mbar_init(&x_taken[b], 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&in_full[b], 1);
      mbar_init(&stg_full[b], kWbMathWarps);       // one arrive per warp: N arrivals on one mbarrier serialise
      mbar_init(&stg_free[b], 1);
    }
    for (int d = 0; d < kWbDeps; ++d) {
      mbar_init(&dep_ready[d], 1);
      mbar_init(&dep_free[d], is_R ? (has_x ? 2 + kWbMathWarps : 2) : 1);   // producer (+ input loader (+ the math warps) of R)
    }
    fence_mbar_init();
    // the first use of both receive buffers is armed here; later uses are armed by math thread 0 BEFORE it tells the
    // senders that the buffer is free, so a complete_tx can never precede its expect_tx
    mbar_expect_tx(&recv_full[0], kWbRecvBytes);
    mbar_expect_tx(&recv_full[1], kWbRecvBytes);
    tma_prefetch_desc(&ly.t_dg);
  }
  if (warp == kWbWarpMma) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  cluster_sync_all();                          // every CTA's barriers are initialised before any remote arrive / push
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  // ---- weights -> tensor memory (once): lane r = unit 128 u + r, K slice [768 s, 768 s + 768) of the 4H gate columns
  if (warp < 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const __nv_bfloat16* wrow = (is_R ? ly.whhT : ly.wihT) + (size_t)(u * 128 + r) * (4 * H) + s * (kWbKb * 64);
    for (int kb = 0; kb < kWbKb; ++kb) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(wrow + kb * 64 + i * 8));
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
      }
      tmem_st32(tmem + (uint32_t(q * 32) << 16) + kb * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kWbWarpPoll && lane == 0) {
    // ------------------------------------------------------------------ dependency poller
    int it = 0;
    for (int t = T - 1; t >= 0; --t) {
      for (int j = 0; j < nt; ++j, ++it) {
        const int d = (int)(it % kWbDeps);
        WB_ACC(0, mbar_wait(&dep_free[d], (uint32_t)(((it / kWbDeps) & 1) ^ 1)));
        if (WB_ABL(8)) {
        } else if (is_R) {
          WB_ACC(1, wait_two_counters(t < T - 1 ? p.dcnt + l * nt + j : nullptr, (unsigned)(NS * (T - 1 - t)),
                            has_x ? p.xcnt + ((size_t)(l + 1) * NS + ns) * nt + j : nullptr, (unsigned)(T - t)));
        } else {
          WB_ACC(1, wait_two_counters(p.dcnt + l * nt + j, (unsigned)(NS * (T - t)),
                            T - t > kWbXRing ? p.dcnt + (l - 1) * nt + j : nullptr, (unsigned)(NS * (T - t - kWbXRing))));
        }
        mbar_arrive(&dep_ready[d]);
      }
    }
    WB_ACC_OUT(0, 2);
  } else if (warp == kWbWarpLoad && lane == 0) {
    // ------------------------------------------------------------------ epilogue-input loader (R only)
    if (is_R) {
      int it = 0;
      for (int t = T - 1; t >= 0; --t) {
        for (int j = 0; j < nt; ++j, ++it) {
          const int buf = (int)(it & 1);
          const int d = (int)(it % kWbDeps);
          WB_ACC(0, mbar_wait(&dep_ready[d], (uint32_t)((it / kWbDeps) & 1)));   // our own dc store of frame t+1 is complete
          mbar_arrive(&dep_free[d]);
          const int b3 = (int)(it % 3);
          const uint32_t par3 = (uint32_t)((it / 3) & 1);
          if (!WB_ABL(2048)) WB_ACC(1, mbar_wait(&stg_free[b3], par3 ^ 1));             // the stores of tile it-3 have left the buffer
          if (it >= 2 && !WB_ABL(2048)) WB_ACC(2, mbar_wait(&stg_full[(it - 2) % 3], (uint32_t)(((it - 2) / 3) & 1)));   // c buffer: math of it-2 done
          uint8_t* sb = stg + b3 * kWbIoBytes;
          uint8_t* cb = cst + buf * kWbCpBytes;
          if (WB_ABL(32)) { mbar_arrive(&in_full[b3]); continue; }
          mbar_expect_tx(&in_full[b3], kWbIoBytes + kWbCpBytes);
          tma_load_3d(sb + kWbOffG, &ly.t_dg, &in_full[b3], ns * 128, j * kWbTile, t);
          tma_load_3d(sb + kWbOffG + 8192, &ly.t_dg, &in_full[b3], ns * 128 + 64, j * kWbTile, t);
          tma_load_3d(cb, &ly.t_c, &in_full[b3], ns * 32, j * kWbTile, t);
          tma_load_3d(sb + kWbOffDc, &ly.t_dc, &in_full[b3], ns * 32, j * kWbTile, 0);
        }
      }
      WB_ACC_OUT(2, 3);
    }
  } else if (warp == kWbWarpTma) {
    // ------------------------------------------------------------------ TMA producer (B operand = dG rows)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = T - 1; t >= 0; --t) {
        for (int j = 0; j < nt; ++j, ++it) {
          const int d = (int)(it % kWbDeps);
          WB_ACC(0, mbar_wait(&dep_ready[d], (uint32_t)((it / kWbDeps) & 1)));
          mbar_arrive(&dep_free[d]);
          const int slab = is_R ? t + 1 : t;
          for (int gi = 0; gi < kWbKb / kWbKbPerStage; ++gi) {
            if (!WB_ABL(32768)) WB_ACC(1, mbar_wait(&empty[stage], phase ^ 1));
            if (WB_ABL(4)) mbar_arrive(&full[stage]); else mbar_expect_tx(&full[stage], kWbStageBytes);
#pragma unroll
            for (int k = 0; k < kWbKbPerStage; ++k)
              if (!WB_ABL(4)) tma_load_3d(ring + stage * kWbStageBytes + k * (kWbTile * 128), &ly.t_dg, &full[stage],
                          s * (kWbKb * 64) + (gi * kWbKbPerStage + k) * 64, j * kWbTile, slab);
            if (++stage == kWbStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      WB_ACC_OUT(5, 2);
    }
  } else if (warp == kWbWarpMma) {
    // ------------------------------------------------------------------ MMA issuer (A = W^T slice in TMEM, bf16)
    constexpr uint32_t idesc = umma_idesc_bf16(128, kWbTile, 0, 0);
    const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(ring));
    const uint32_t desc_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
    int stage = 0;
    uint32_t phase = 0;
    // `if (elect_one())` around the loop, NOT a per-thread trip count: with `n = elect_one() ? total : 0; for (it < n)`
    // the compiler treats the body as divergent code and wraps EVERY tcgen05.mma in an elect / R2UR.BROADCAST /
    // BRA.U.ANY loop (~12 instructions, 60-70 cycles of issue per MMA instead of 32).
    if (elect_one())
    for (int it = 0; it < total; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t upar = (uint32_t)((it >> 1) & 1);
      long long* tr = (trace_cta && (T - 1 - it / nt) == T / 2) ? trace_cta + (it % nt) * 16 : nullptr;
      WL_STAMP(0);
      if (!WB_ABL(16384)) WB_ACC(0, mbar_wait(&acc_empty[buf], upar ^ 1));
      tc_fence_after();
      for (int gi = 0; gi < kWbKb / kWbKbPerStage; ++gi) {
        if (!WB_ABL(32768)) WB_ACC(1, mbar_wait(&full[stage], phase));
        tc_fence_after();
        if (gi == kWbKb / kWbKbPerStage - 1) WL_STAMP(3);
        const uint32_t lo = desc_lo0 + stage * (kWbStageBytes >> 4);
        const uint32_t a0 = tmem + gi * kWbKbPerStage * 32;
        const uint32_t dacc = tmem + kWbAccCol + buf * kWbTile;
#pragma unroll
        for (int q = 0; q < 4 * kWbKbPerStage; ++q)
          if (!WB_ABL(1)) umma_f16_ts_lohi(dacc, a0 + q * 8, lo + (q >> 2) * ((kWbTile * 128) >> 4) + (q & 3) * 2, desc_hi, idesc,
                           q == 0 ? (gi != 0 ? 1u : 0u) : 1u);
        umma_commit(&empty[stage]);
        if (gi == kWbKb / kWbKbPerStage - 1) umma_commit(&acc_full[buf]);
        if (++stage == kWbStages) { stage = 0; phase ^= 1; }
      }
#ifdef SVB_WB_ACCOUNT
      if (it == total - 1) WB_ACC_OUT(7, 2);
#endif
    }
  } else if (warp >= kWbWarpXch && warp < kWbWarpXch + 4) {
    // ------------------------------------------------------------------ exchange warps: TMEM -> staged quarters
    // One warp per TMEM lane quarter (= unit quarter of the tile), 64 rows in two halves.  The quarter this CTA owns
    // (q == s) is staged as fp32 for the math warps; the other three are staged as fp16 and pushed to their owner by
    // lane 0 of the staging warp itself, half by half (2 KB bulk copies: the first half travels while the second is
    // staged; no push thread, nobody waits for the slowest warp).
    const int q = warp & 3;
    const bool owner = q == s;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    const uint32_t dslot = (uint32_t)((q - s - 1) & 3);          // index of the quarter in my send buffer
    const uint32_t src0 = owner ? smem_u32(own) : smem_u32(send) + dslot * kWbQuarterBytes;
    // receiver q keeps my quarter in slot (s - q - 1) mod 4 of its receive buffer
    const uint32_t rdst0 = map_to_cta(smem_u32(recv) + (uint32_t)((s - q - 1) & 3) * kWbQuarterBytes, (uint32_t)q);
    const uint32_t rbar0 = map_to_cta(smem_u32(recv_full), (uint32_t)q);
    for (int it = 0; it < total; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t upar = (uint32_t)((it >> 1) & 1);
      // (debug stamps: the owner warp and the sender warp of the next quarter)
      long long* tr = (trace_cta && (T - 1 - it / nt) == T / 2 && lane == 0 && (q == s || q == ((s + 1) & 3))) ? trace_cta + (it % nt) * 16 : nullptr;
      if (owner) WL_STAMP(4);
      if (WB_ABL(512)) {
      } else if (owner) WB_ACC(0, mbar_wait_warp(&consumed[buf], upar ^ 1));            // the math warps have read own[buf] of use k-1
      else if (!WB_ABL(64)) WB_ACC(0, mbar_wait_cluster_warp(&peer_free[buf], upar ^ 1));         // send[buf] has left, the receiver is free and armed
      if (owner) WL_STAMP(5); else WL_STAMP(1);
      if (!WB_ABL(65536)) WB_ACC(1, mbar_wait_warp(&acc_full[buf], upar));
      tc_fence_after();
      if (owner) WL_STAMP(6);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(tmem + lane_base + kWbAccCol + buf * kWbTile + half * 32, v);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (WB_ABL(16)) {
        } else if (owner) {
          WbStage<0>::own(src0 + buf * kWbOwnBytes + half * 32 * 128 + lane * 4, v);
        } else {
          WbStage<0>::send(src0 + buf * kWbSendBytes + half * 32 * 64 + lane * 2, v);
        }
        if (!owner) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !WB_ABL(64))
            dsmem_bulk_push(rdst0 + buf * kWbRecvBytes + half * (kWbQuarterBytes / 2),
                            src0 + buf * kWbSendBytes + half * (kWbQuarterBytes / 2), kWbQuarterBytes / 2, rbar0 + buf * 8);
        }
      }
      if (owner) {                              // own quarter staged: counts into the same barrier as the three pushes
        __syncwarp();
        if (lane == 0) mbar_arrive(&recv_full[buf]);
      }
      if (owner) WL_STAMP(7); else WL_STAMP(2);
    }
    // quarter 0 and quarter 1 warps: one of them is the owner (consumed wait), the other a sender (peer_free wait)
    if (lane == 0 && q == 0) WB_ACC_OUT(9, 2);
    if (lane == 0 && q == 1) WB_ACC_OUT(12, 2);
  } else if (warp < kWbMathWarps) {
    // ------------------------------------------------------------------ math warps: split-K reduction (own fp32 quarter
    // + three received fp16 quarters + dX of the layer above), then X: dX tile out; R: gate backward, in place.
    // 16 warps, thread = (row, 4 units): with 8 warps of (row, 8 units) the two warps per scheduler could not hide their
    // own latencies (3100 busy cycles per tile for ~420 instructions per thread).
    const uint32_t stg_a = smem_u32(stg), own_a = smem_u32(own), recv_a = smem_u32(recv);
    const int mt = threadIdx.x;                // 0..511
    const int row = mt >> 3, ug = mt & 7;      // units 4 ug .. 4 ug + 3 of the CTA's 32
    float bsum[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
    // Software pipeline (R, nt >= 2): step k reduces tile k (waits for the exchange of tile k, sends the credits)
    // and THEN runs the gate backward of tile k - 1, so that the exchange loop (stage -> push -> reduce -> credit ->
    // stage of the tile after next) does not contain the gate math.  With one tile per frame tile k's operands are the
    // dG of tile k - 1: no look-ahead (lag 0).
    const int lag = (is_R && nt >= 2) ? 1 : 0;
    float dh[4] = {0.f, 0.f, 0.f, 0.f};          // reduced dL/dh of the tile whose gate backward is next
    int t = T - 1, j = 0;                        // tile being reduced
    int tg = T - 1, jg = 0;                      // tile whose gates are processed (lag tiles behind)
    // per-tile bookkeeping kept short (it sits on the critical loop of the kernel, section 4.2 of DESIGN.md): offsets
    // that depend on the thread only are formed once, ring slot / buffer indices are carried instead of divided out
    const uint32_t own_t = own_a + row * 128 + ug * 16, recv_t = recv_a + row * 64 + ug * 8;
    const size_t x_t = ((size_t)ns * kWbTile + row) * 32 + ug * 4;          // floats inside a (frame slot, tile) block
    const size_t x_tile = (size_t)NS * kWbTile * 32;                         // floats per (frame slot, tile)
    int xslot = (T - 1) % kWbXRing;                                          // t % kWbXRing, carried
    int b3g = 0;                                                             // ig % 3 and (ig / 3) & 1, carried
    uint32_t par3g = 0;
    const int G_ = ug >> 1;
    const uint32_t g_off_t = kWbOffG + (G_ >> 1) * 8192 + row * 128 + 8 * (ug & 1);
    const uint32_t u_off_t = row * 128 + ((ug ^ (row & 7)) << 4);
    uint32_t gsw[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) gsw[g] = g_off_t + (((4 * (G_ & 1) + g) ^ (row & 7)) << 4);
    for (int it = 0; it < total + lag; ++it) {
      float dn[4] = {0.f, 0.f, 0.f, 0.f};
      if (it < total) {
        const int buf = (int)(it & 1);
        const uint32_t upar = (uint32_t)((it >> 1) & 1);
        long long* tr = (trace_cta && t == T / 2 && threadIdx.x == 0) ? trace_cta + j * 16 : nullptr;
        WL_STAMP(8);
        float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* xrow = nullptr;
        if (has_x) {
          // dX tile of the layer above (L2-resident ring), in flight while the partial sums arrive
          const int d = (int)(it % kWbDeps);
          WB_ACC(0, mbar_wait_warp(&dep_ready[d], (uint32_t)((it / kWbDeps) & 1)));
          const float* xq = p.layer[l + 1].xring + (size_t)(xslot * nt + j) * x_tile + x_t;
          xrow = xq - ug * 4;
          x0 = __ldcg(reinterpret_cast<const float4*>(xq));
          __syncwarp();
          if (lane == 0) mbar_arrive(&dep_free[d]);
        } else if (is_R && t == T - 1 && j * kWbTile + row < p.B) {       // top layer, last frame: + dL/dh_last
          x0 = __ldg(reinterpret_cast<const float4*>(p.dh_last + (size_t)(j * kWbTile + row) * H + ns * 32) + ug);
        }
        if (!WB_ABL(64)) WB_ACC(2, mbar_wait_warp(&recv_full[buf], upar));   // own quarter + the three received quarters
        const float4 o0 = lds_f4(own_t + buf * kWbOwnBytes);
        WL_STAMP(9);
        dn[0] = o0.x; dn[1] = o0.y; dn[2] = o0.z; dn[3] = o0.w;
#pragma unroll
        for (int sl = 0; sl < 3; ++sl) {
          const uint2 r = lds_u2(recv_t + buf * kWbRecvBytes + sl * kWbQuarterBytes);
          const float2 a0 = half2_to_float2(r.x), a1 = half2_to_float2(r.y);
          dn[0] += a0.x; dn[1] += a0.y; dn[2] += a1.x; dn[3] += a1.y;
        }
        dn[0] += x0.x; dn[1] += x0.y; dn[2] += x0.z; dn[3] += x0.w;
        // every math thread has its part of recv[buf] / own[buf] in registers: re-arm the receive barrier for the use
        // after next, hand own[buf] back to the exchange warps, tell the three senders
        if (!WB_ABL(131072)) WB_ACC(3, math_warps_sync());
        if (warp == 0) {
          if (lane == 0) {
            if (it + 2 < total && !WB_ABL(64)) mbar_expect_tx(&recv_full[buf], kWbRecvBytes);
            mbar_arrive(&consumed[buf]);
          }
          __syncwarp();
          if (lane < 3 && it + 2 < total && !WB_ABL(64)) mbar_arrive_remote_relaxed(&peer_free[buf], (uint32_t)((s + 1 + lane) & 3));
        }
        WL_STAMP(10);
#ifndef SVB_NO_RING_DISCARD
        // the dX row has been consumed: drop its line from L2 without write-back (the ring is read exactly once, from L2)
        if (has_x && ug == 0) asm volatile("discard.global.L2 [%0], 128;" ::"l"(xrow) : "memory");
#endif
        if (!is_R) {
          // dX tile [row][32 units] fp32: staged in shared memory (the epilogue staging area is unused in X) and
          // bulk-copied to the ring by the signal thread
          WB_ACC(4, mbar_wait_warp(&x_taken[buf], upar ^ 1));    // the copy of tile it-2 has left the buffer
          sts_f4(stg_a + buf * 8192 + row * 128 + ug * 16, make_float4(dn[0], dn[1], dn[2], dn[3]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_done[buf]);
        }
        if (++j == nt) { j = 0; --t; if (--xslot < 0) xslot = kWbXRing - 1; }
      }
      if (lag == 0) { dh[0] = dn[0]; dh[1] = dn[1]; dh[2] = dn[2]; dh[3] = dn[3]; }
      if (is_R && it >= lag) {
        const int ig = it - lag;                   // tile (tg, jg)
        const int bufg = (int)(ig & 1);
        long long* tr = (trace_cta && tg == T / 2 && threadIdx.x == 0) ? trace_cta + jg * 16 : nullptr;
        const int b3 = b3g;
        const uint32_t par3 = par3g;
        if (++b3g == 3) { b3g = 0; par3g ^= 1; }
        if (!WB_ABL(1024)) WB_ACC(4, mbar_wait_warp(&in_full[b3], par3));
        WL_STAMP(11);
        const uint32_t sb = stg_a + b3 * kWbIoBytes;
        const uint32_t cb = stg_a + kWbIo * kWbIoBytes + bufg * kWbCpBytes;
        if (!WB_ABL(2)) {
          const uint32_t u_off = u_off_t;
          const float4 cp4 = lds_f4(cb + u_off), dc4 = lds_f4(sb + kWbOffDc + u_off);
          uint32_t gaddr[4];
          uint2 gq[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            gaddr[g] = sb + gsw[g];
            gq[g] = lds_u2(gaddr[g]);
          }
          const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, dcs[4] = {dc4.x, dc4.y, dc4.z, dc4.w};
          float gi[4], gf[4], gg[4], go[4], di[4], df[4], dg[4], dO[4], dcn[4];
          // the stash holds the activations as fp16 (pack8_stash); dG goes back in place as bf16
          { const float2 a = half2_to_float2(gq[0].x), b = half2_to_float2(gq[0].y); gi[0] = a.x; gi[1] = a.y; gi[2] = b.x; gi[3] = b.y; }
          { const float2 a = half2_to_float2(gq[1].x), b = half2_to_float2(gq[1].y); gf[0] = a.x; gf[1] = a.y; gf[2] = b.x; gf[3] = b.y; }
          { const float2 a = half2_to_float2(gq[2].x), b = half2_to_float2(gq[2].y); gg[0] = a.x; gg[1] = a.y; gg[2] = b.x; gg[3] = b.y; }
          { const float2 a = half2_to_float2(gq[3].x), b = half2_to_float2(gq[3].y); go[0] = a.x; go[1] = a.y; go[2] = b.x; go[3] = b.y; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float tc = tanh_approx(fmaf(gf[i], cp[i], gi[i] * gg[i]));     // tanh(c_t), c_t = f c_{t-1} + i g
            const float dc = dh[i] * go[i] * (1.f - tc * tc) + dcs[i];
            dO[i] = dh[i] * tc * go[i] * (1.f - go[i]);
            di[i] = dc * gg[i] * gi[i] * (1.f - gi[i]);
            df[i] = dc * cp[i] * gf[i] * (1.f - gf[i]);
            dg[i] = dc * gi[i] * (1.f - gg[i] * gg[i]);
            dcn[i] = dc * gf[i];
          }
          // bias gradients: per-thread partial column sums of dG over all tiles and frames (reduced over the 64 rows
          // once, at the end of the kernel)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            bsum[i] += di[i]; bsum[4 + i] += df[i]; bsum[8 + i] += dg[i]; bsum[12 + i] += dO[i];
          }
          sts_u2(gaddr[0], make_uint2(pack_bf16x2(di[0], di[1]), pack_bf16x2(di[2], di[3])));
          sts_u2(gaddr[1], make_uint2(pack_bf16x2(df[0], df[1]), pack_bf16x2(df[2], df[3])));
          sts_u2(gaddr[2], make_uint2(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3])));
          sts_u2(gaddr[3], make_uint2(pack_bf16x2(dO[0], dO[1]), pack_bf16x2(dO[2], dO[3])));
          sts_f4(sb + kWbOffDc + u_off, make_float4(dcn[0], dcn[1], dcn[2], dcn[3]));
        }
        if (!WB_ABL(256)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_full[b3]);
        WL_STAMP(14);
        if (++jg == nt) { jg = 0; --tg; }
      }
      dh[0] = dn[0]; dh[1] = dn[1]; dh[2] = dn[2]; dh[3] = dn[3];
    }
    if (threadIdx.x == 0) WB_ACC_OUT(14, 5);
    if (is_R) {
      // partial sums -> shared memory [row][packed column] (the operand ring is idle by now: the last MMA of this CTA
      // completed before the last accumulator was handed to the exchange warps)
      float* bsc = reinterpret_cast<float*>(ring);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i) bsc[row * 128 + 32 * (ug >> 1) + 8 * g + 4 * (ug & 1) + i] = bsum[g * 4 + i];
    }
  } else if (warp == kWbWarpStore && !is_R) {
    // ------------------------------------------------------------------ signal thread (X): publish dX tiles (lazily,
    // two tiles in flight: see the R store thread below)
    if (lane == 0) {
      bool pend = false;
      unsigned* pflag = nullptr;
      unsigned pval = 0;
      int t = T - 1, j = 0;
      for (int it = 0; it < total; ++it) {
        const int buf = (int)(it & 1);
        const uint32_t upar = (uint32_t)((it >> 1) & 1);
        if (pend && !mbar_test_wait(&x_done[buf], upar)) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("red.release.gpu.global.max.u32 [%0], %1;" ::"l"(pflag), "r"(pval) : "memory");
          pend = false;
        }
        mbar_wait(&x_done[buf], upar);
        float* dst = ly.xring + (((size_t)((t % kWbXRing) * nt + j) * NS + ns) * 512) * 4;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg) + buf * 8192),
                     "n"(8192) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&x_taken[buf]);
        if (pend) {
          asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
          asm volatile("red.release.gpu.global.max.u32 [%0], %1;" ::"l"(pflag), "r"(pval) : "memory");
        }
        pend = true;
        pflag = p.xcnt + ((size_t)l * NS + ns) * nt + j;
        pval = (unsigned)(T - t);
        if (++j == nt) { j = 0; --t; }
      }
      if (pend) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("red.release.gpu.global.max.u32 [%0], %1;" ::"l"(pflag), "r"(pval) : "memory");
      }
    }
  } else if (warp == kWbWarpStore && is_R) {
    // ------------------------------------------------------------------ store + signal thread (R)
    // The completion of a tile's TMA stores (cp.async.bulk.wait_group, needed before the release that publishes the
    // tile) takes 3400-4700 cycles under load -- waited for tile by tile it was as long as the kernel's period.  A
    // tile is therefore published LAZILY: if the next tile is already staged its stores are issued first and the
    // thread then waits for the older group only (wait_group 1: two tiles in flight); if the next tile is not ready
    // the pending one is completed and published at once, so a release never waits for a later tile (with one tile per
    // frame the next frame cannot start before it).  The in-place buffer is handed back as soon as it has been read.
    if (elect_one()) {
      bool pend = false;
      unsigned* pflag = nullptr;
      long long* ptr_tr = nullptr;
      int it = 0;
      for (int t = T - 1; t >= 0; --t) {
        for (int j = 0; j < nt; ++j, ++it) {
          const int b3 = (int)(it % kWbIo);
          const uint32_t par3 = (uint32_t)((it / kWbIo) & 1);
          if (WB_ABL(4096)) pend = false;
          if (pend && !mbar_test_wait(&stg_full[b3], par3)) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            red_release_add(pflag, 1u);      // release: see wlstm.cuh
            if (ptr_tr) *ptr_tr = clock64();
            pend = false;
          }
          WB_ACC(0, mbar_wait(&stg_full[b3], par3));
          const uint8_t* sb = stg + b3 * kWbIoBytes;
          if (!WB_ABL(128)) {
            tma_store_3d(&ly.t_dg, sb + kWbOffG, ns * 128, j * kWbTile, t);
            tma_store_3d(&ly.t_dg, sb + kWbOffG + 8192, ns * 128 + 64, j * kWbTile, t);
            tma_store_3d(&ly.t_dc, sb + kWbOffDc, ns * 32, j * kWbTile, 0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          WB_ACC(1, asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"));
          mbar_arrive(&stg_free[b3]);
          if (pend && !WB_ABL(4096)) {
            WB_ACC(2, asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"));
            red_release_add(pflag, 1u);
            if (ptr_tr) *ptr_tr = clock64();
          }
          pend = true;
          pflag = p.dcnt + l * nt + j;
          ptr_tr = (trace_cta && t == T / 2) ? trace_cta + j * 16 + 15 : nullptr;
        }
      }
      if (pend) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        red_release_add(pflag, 1u);
        if (ptr_tr) *ptr_tr = clock64();
      }
      WB_ACC_OUT(19, 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (is_R && threadIdx.x < 128 && ly.gbias_ih) {
    // this CTA owns packed gate columns [128 ns, 128 ns + 128) for all rows and frames: finish the bias gradients
    const float* bsc = reinterpret_cast<const float*>(ring);
    float sum = 0.f;
    for (int r64 = 0; r64 < 64; ++r64) sum += bsc[r64 * 128 + threadIdx.x];
    const int pc = ns * 128 + threadIdx.x;                           // packed column
    const int r = ((pc & 31) >> 3) * H + (pc >> 5) * 8 + (pc & 7);   // reference row (gate-major i|f|g|o)
    sum *= __ldg(p.inv_scale);
    ly.gbias_ih[r] = sum;
    ly.gbias_hh[r] = sum;
  }
  cluster_sync_all();                          // no CTA leaves while a peer may still push into or signal it
  if (warp == kWbWarpMma) tmem_dealloc<512>(tmem);
  if (p.trace && threadIdx.x == 0) {
    p.trace[2 * nt * 16 + blockIdx.x] = clock64() - t_cta0;
#ifdef SVB_WB_ACCOUNT
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[2 * nt * 16 + 256 + 32 * blockIdx.x + 31] = smid;
#endif
  }
}

// Largest number of 4-CTA clusters of the kernel that can be co-resident on the current device (0 on error).
template <int H>
static int wbptt_max_clusters() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device_index();
  if (!cached[dev]) {
    auto kern = wbptt_kernel<H>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWbSmem) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(kWbThreads);
    cfg.dynamicSmemBytes = kWbSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cached[dev] = n > 0 ? n : -1;
  }
  return cached[dev] > 0 ? cached[dev] : 0;
}

template <int H>
static int launch_wbptt(WbParams& p, cudaStream_t s) {
  auto kern = wbptt_kernel<H>;
  static unsigned long long configured = 0;        // one bit per device
  if (first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWbSmem);
    if (e != cudaSuccess) { set_error("wbptt: cudaFuncSetAttribute", e); return SVB_ERR_CUDA; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * (2 * p.L - 1) * (H / 128));
  cfg.blockDim = dim3(kWbThreads);
  cfg.dynamicSmemBytes = kWbSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // Cooperative cluster launch: the driver guarantees that all 30 clusters are co-resident (the kernel's flag protocol
  // spins on other CTAs' progress), whatever else is running -- NCCL kernels of an overlapped all-reduce, a second
  // stream.  Nsight Compute refuses the cooperative + cluster combination (LaunchFailed), so profiling runs set
  // SVB_PLAIN_CLUSTER_LAUNCH=1: a plain cluster launch, safe only while nothing else holds SMs (one 640-thread, 218 KB
  // CTA per SM; the occupancy query in svb_embedder_backward has checked that the clusters fit an empty device).
  static const bool plain = getenv("SVB_PLAIN_CLUSTER_LAUNCH") != nullptr;
  cfg.numAttrs = plain ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) { set_error("wbptt: cluster launch", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
