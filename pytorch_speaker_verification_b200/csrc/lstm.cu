// SpeechEmbedder forward / backward for sm_100a: 3-layer LSTM (tcgen05 GEMMs) + last-frame projection + L2 norm.
//
// Replaces speech_embedder_net.py:27-33 (SpeechEmbedder.forward: nn.LSTM -> cuDNN/oneDNN, nn.Linear, torch.norm)
// and the BPTT that autograd runs for train_speech_embedder.py:62 in the reference.
//
// Data layout (all device memory, allocated by the caller):
//   x_tm      [T, B, Ip]   bf16 hi / lo   time-major copy of the (B, T, I) input, Ip = I rounded up to 8
//   gin       [T, B, 4H]   fp32           input projection W_ih x_t + b_ih + b_hh, PACKED gate columns
//   hseq[l]   [T+1, B, H]  bf16 hi / lo   h_t in slot t+1, slot 0 = h_{-1} = 0
//   cseq[l]   [T+1, B, H]  fp32           c_t in slot t+1 (training; 2 slots otherwise)
//   gates[l]  [T+1, B, 4H] bf16           sigma/tanh gate activations (training), overwritten in place by the
//                                         gate pre-activation gradients dG during BPTT; slot T = 0
// Packed gate order: column p = 32*(u/8) + 8*g + (u%8) for gate g in (i,f,g,o) of hidden unit u, so every
// 32-column accumulator chunk that an epilogue thread owns holds all four gates of 8 units and the cell update
// is fused into the recurrent GEMM's epilogue with no exchange between threads.
// Numerics: recurrent GEMM bf16 x bf16 -> fp32; input projection split-bf16, 3 terms (x_hi W_hi + x_lo W_hi +
// x_hi W_lo) accumulated in one TMEM accumulator; gates, cell state, projection and norm in fp32.
#include "tc_gemm.cuh"
#include "epilogues.cuh"
#include "../../include/svb200.h"

namespace svb {
void set_error(const char* what, cudaError_t e);
int make_operand_map(CUtensorMap* out, const void* p, int rows, int K, int64_t ld, int mn_major, int box_rows_kmajor);

__host__ __device__ inline int packed_col(int g, int u) { return 32 * (u >> 3) + 8 * g + (u & 7); }
static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }
static inline int round8(int x) { return (x + 7) & ~7; }

// ------------------------------------------------------------------------------------------ packed weights
struct LayerW {
  __nv_bfloat16 *wih_hi, *wih_lo;   // [4H, Ip] packed rows
  __nv_bfloat16 *whh_hi, *whh_lo;   // [4H, H]  packed rows
  float* bias;                      // [4H] packed, b_ih + b_hh
  int I, Ip;
};
struct PackedW {
  LayerW l[8];
  size_t bytes;
};
static PackedW layout_packed(char* base, int I, int H, int L) {
  PackedW pw{};
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al256(bytes); return p; };
  for (int l = 0; l < L; ++l) {
    const int Il = l == 0 ? I : H, Ip = round8(Il);
    pw.l[l].I = Il; pw.l[l].Ip = Ip;
    pw.l[l].wih_hi = (__nv_bfloat16*)take((size_t)4 * H * Ip * 2);
    pw.l[l].wih_lo = (__nv_bfloat16*)take((size_t)4 * H * Ip * 2);
    pw.l[l].whh_hi = (__nv_bfloat16*)take((size_t)4 * H * H * 2);
    pw.l[l].whh_lo = (__nv_bfloat16*)take((size_t)4 * H * H * 2);
    pw.l[l].bias = (float*)take((size_t)4 * H * 4);
  }
  pw.bytes = off;
  return pw;
}

__global__ void pack_weights_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, LayerW lw, int H) {
  const int p = blockIdx.x;                       // packed row
  const int g = (p & 31) >> 3, u = (p >> 5) * 8 + (p & 7);
  const int r = g * H + u;                        // reference row (gate-major: i|f|g|o, speech_embedder_net.py:19)
  for (int k = threadIdx.x; k < lw.Ip; k += blockDim.x) {
    const float v = k < lw.I ? w_ih[(size_t)r * lw.I + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    lw.wih_hi[(size_t)p * lw.Ip + k] = hi;
    lw.wih_lo[(size_t)p * lw.Ip + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float v = w_hh[(size_t)r * H + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    lw.whh_hi[(size_t)p * H + k] = hi;
    lw.whh_lo[(size_t)p * H + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  if (threadIdx.x == 0) lw.bias[p] = b_ih[r] + b_hh[r];
}

// ------------------------------------------------------------------------------------------ workspace
struct Dims { int B, T, I, H, L, P; };
struct Work {
  __nv_bfloat16 *x_hi, *x_lo;       // [T, B, Ip0]
  float* gin;                       // [T, B, 4H]
  __nv_bfloat16 *h_hi[8], *h_lo[8]; // [T+1, B, H]
  float* c[8];                      // [cslots, B, H]
  __nv_bfloat16* gates[8];          // [T+1, B, 4H] (training)
  float *h_last, *y, *inv_norm;     // [B,H], [B,P], [B]
  float *dh_above, *dc, *dh_last, *dy;  // backward: [T,B,H], [B,H], [B,H], [B,P]
  float* colsum_part;               // [chunks, 4H]
  int cslots;
  size_t bytes;
};
constexpr int kColsumRows = 2048;
static Work layout_work(char* base, const Dims& d, int training) {
  Work w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al256(bytes); return p; };
  const size_t B = d.B, T = d.T, H = d.H, Ip0 = round8(d.I);
  w.cslots = training ? d.T + 1 : 2;
  w.x_hi = (__nv_bfloat16*)take(T * B * Ip0 * 2);
  w.x_lo = (__nv_bfloat16*)take(T * B * Ip0 * 2);
  w.gin = (float*)take(T * B * 4 * H * 4);
  for (int l = 0; l < d.L; ++l) {
    w.h_hi[l] = (__nv_bfloat16*)take((T + 1) * B * H * 2);
    w.h_lo[l] = (__nv_bfloat16*)take((T + 1) * B * H * 2);
    w.c[l] = (float*)take((size_t)w.cslots * B * H * 4);
    w.gates[l] = training ? (__nv_bfloat16*)take((T + 1) * B * 4 * H * 2) : nullptr;
  }
  w.h_last = (float*)take(B * H * 4);
  w.y = (float*)take(B * d.P * 4);
  w.inv_norm = (float*)take(B * 4);
  if (training) {
    w.dh_above = (float*)take(T * B * H * 4);
    w.dc = (float*)take(B * H * 4);
    w.dh_last = (float*)take(B * H * 4);
    w.dy = (float*)take(B * d.P * 4);
    w.colsum_part = (float*)take(((T * B + kColsumRows - 1) / kColsumRows) * 4 * H * 4);
  }
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------------------ input prep
template <typename Tin>
__global__ void prep_x_kernel(const Tin* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                              int B, int T, int I, int Ip) {
  const size_t n = (size_t)T * B * Ip;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % Ip;
    const size_t tb = i / Ip;
    const int b = tb % B, t = tb / B;
    const float v = k < I ? (float)x[((size_t)b * T + t) * I + k] : 0.f;   // x.float() (speech_embedder_net.py:28)
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------------------------------ LSTM epilogues
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_of(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi_of(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Forward cell: acc = W_hh h_{t-1} for 8 units x 4 gates (packed chunk); adds gin, applies the gates.
struct EpiLstmFwd {
  struct Params {
    const float* gin;          // [B, 4H] slice of step t
    const float* c_prev;       // [B, H]
    float* c_out;              // [B, H]
    __nv_bfloat16* h_hi;       // [B, H] slot t+1
    __nv_bfloat16* h_lo;       // nullable
    __nv_bfloat16* gates;      // [B, 4H] slice of step t, nullable
    float* h_f32;              // [B, H], nullable (top layer, last step)
    int H;
  };
  struct Tile {};
  static __device__ __forceinline__ void prologue(const Params& p, Tile&, int m, int n0, bool valid) {
    if (valid) {
      const float* g = p.gin + (size_t)m * 4 * p.H + n0;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 32));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 64));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 96));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.c_prev + (size_t)m * p.H + (n0 >> 2)));
    }
  }
  static __device__ __forceinline__ void apply(const Params& p, Tile&, int m, int n0, float (&acc)[32]) {
    const int u0 = (n0 >> 5) * 8;
    const float4* g4 = reinterpret_cast<const float4*>(p.gin + (size_t)m * 4 * p.H + n0);
    float pre[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 v = __ldg(g4 + j);
      pre[4 * j + 0] = acc[4 * j + 0] + v.x; pre[4 * j + 1] = acc[4 * j + 1] + v.y;
      pre[4 * j + 2] = acc[4 * j + 2] + v.z; pre[4 * j + 3] = acc[4 * j + 3] + v.w;
    }
    const size_t hoff = (size_t)m * p.H + u0;
    const float4 c0 = *reinterpret_cast<const float4*>(p.c_prev + hoff);
    const float4 c1 = *reinterpret_cast<const float4*>(p.c_prev + hoff + 4);
    const float cp[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float cn[8], hn[8], gi[8], gf[8], gg[8], go[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gi[j] = sigmoidf_fast(pre[j]);
      gf[j] = sigmoidf_fast(pre[8 + j]);
      gg[j] = tanhf_fast(pre[16 + j]);
      go[j] = sigmoidf_fast(pre[24 + j]);
      cn[j] = gf[j] * cp[j] + gi[j] * gg[j];
      hn[j] = go[j] * tanhf_fast(cn[j]);
    }
    *reinterpret_cast<float4*>(p.c_out + hoff) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    *reinterpret_cast<float4*>(p.c_out + hoff + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
    uint32_t hh[4], hl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hh[j] = pack_bf16x2(hn[2 * j], hn[2 * j + 1]);
      hl[j] = pack_bf16x2(hn[2 * j] - bf16_lo_of(hh[j]), hn[2 * j + 1] - bf16_hi_of(hh[j]));
    }
    *reinterpret_cast<uint4*>(p.h_hi + hoff) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    if (p.h_lo) *reinterpret_cast<uint4*>(p.h_lo + hoff) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
    if (p.h_f32) {
      *reinterpret_cast<float4*>(p.h_f32 + hoff) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      *reinterpret_cast<float4*>(p.h_f32 + hoff + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
    }
    if (p.gates) {
      uint4* go4 = reinterpret_cast<uint4*>(p.gates + (size_t)m * 4 * p.H + n0);
      go4[0] = make_uint4(pack_bf16x2(gi[0], gi[1]), pack_bf16x2(gi[2], gi[3]), pack_bf16x2(gi[4], gi[5]), pack_bf16x2(gi[6], gi[7]));
      go4[1] = make_uint4(pack_bf16x2(gf[0], gf[1]), pack_bf16x2(gf[2], gf[3]), pack_bf16x2(gf[4], gf[5]), pack_bf16x2(gf[6], gf[7]));
      go4[2] = make_uint4(pack_bf16x2(gg[0], gg[1]), pack_bf16x2(gg[2], gg[3]), pack_bf16x2(gg[4], gg[5]), pack_bf16x2(gg[6], gg[7]));
      go4[3] = make_uint4(pack_bf16x2(go[0], go[1]), pack_bf16x2(go[2], go[3]), pack_bf16x2(go[4], go[5]), pack_bf16x2(go[6], go[7]));
    }
  }
};

// Backward cell: acc = dG_{t+1} W_hh (recurrent part of dL/dh_t) for 32 hidden units of one batch row.
struct EpiLstmBwd {
  struct Params {
    __nv_bfloat16* gates;      // [B, 4H] slice of step t: activations in, dG out (in place)
    const float* c_t;          // [B, H]
    const float* c_prev;       // [B, H]
    float* dc;                 // [B, H] running dL/dc (in place)
    const float* dh_above;     // [B, H] gradient from the layer above at step t, nullable
    int H;
    int use_acc;               // 0 at t = T-1 (no recurrent contribution yet)
  };
  struct Tile {};
  static __device__ __forceinline__ void prologue(const Params& p, Tile&, int m, int n0, bool valid) {
    if (valid) {
      const __nv_bfloat16* g = p.gates + (size_t)m * 4 * p.H + 4 * n0;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 64));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.c_t + (size_t)m * p.H + n0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.c_prev + (size_t)m * p.H + n0));
    }
  }
  static __device__ __forceinline__ void apply(const Params& p, Tile&, int m, int n0, float (&acc)[32]) {
    // n0 = first hidden unit of the chunk (multiple of 32) -> packed columns [4 n0, 4 n0 + 128)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t hoff = (size_t)m * p.H + n0 + 8 * q;
      uint4* gp = reinterpret_cast<uint4*>(p.gates + (size_t)m * 4 * p.H + 4 * n0 + 32 * q);
      const uint4 vi = gp[0], vf = gp[1], vg = gp[2], vo = gp[3];
      const uint32_t wi[4] = {vi.x, vi.y, vi.z, vi.w}, wf[4] = {vf.x, vf.y, vf.z, vf.w};
      const uint32_t wg[4] = {vg.x, vg.y, vg.z, vg.w}, wo[4] = {vo.x, vo.y, vo.z, vo.w};
      float ct[8], cp[8], dcs[8], dha[8];
      {
        const float4 a0 = *reinterpret_cast<const float4*>(p.c_t + hoff), a1 = *reinterpret_cast<const float4*>(p.c_t + hoff + 4);
        ct[0] = a0.x; ct[1] = a0.y; ct[2] = a0.z; ct[3] = a0.w; ct[4] = a1.x; ct[5] = a1.y; ct[6] = a1.z; ct[7] = a1.w;
        const float4 b0 = *reinterpret_cast<const float4*>(p.c_prev + hoff), b1 = *reinterpret_cast<const float4*>(p.c_prev + hoff + 4);
        cp[0] = b0.x; cp[1] = b0.y; cp[2] = b0.z; cp[3] = b0.w; cp[4] = b1.x; cp[5] = b1.y; cp[6] = b1.z; cp[7] = b1.w;
        const float4 d0 = *reinterpret_cast<const float4*>(p.dc + hoff), d1 = *reinterpret_cast<const float4*>(p.dc + hoff + 4);
        dcs[0] = d0.x; dcs[1] = d0.y; dcs[2] = d0.z; dcs[3] = d0.w; dcs[4] = d1.x; dcs[5] = d1.y; dcs[6] = d1.z; dcs[7] = d1.w;
        if (p.dh_above) {
          const float4 e0 = *reinterpret_cast<const float4*>(p.dh_above + hoff), e1 = *reinterpret_cast<const float4*>(p.dh_above + hoff + 4);
          dha[0] = e0.x; dha[1] = e0.y; dha[2] = e0.z; dha[3] = e0.w; dha[4] = e1.x; dha[5] = e1.y; dha[6] = e1.z; dha[7] = e1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dha[j] = 0.f;
        }
      }
      float di[8], df[8], dg[8], dO[8], dcn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t si = wi[j >> 1], sf = wf[j >> 1], sg = wg[j >> 1], so = wo[j >> 1];
        const float gi = (j & 1) ? bf16_hi_of(si) : bf16_lo_of(si);
        const float gf = (j & 1) ? bf16_hi_of(sf) : bf16_lo_of(sf);
        const float gg = (j & 1) ? bf16_hi_of(sg) : bf16_lo_of(sg);
        const float go = (j & 1) ? bf16_hi_of(so) : bf16_lo_of(so);
        const float dh = (p.use_acc ? acc[8 * q + j] : 0.f) + dha[j];
        const float tc = tanhf_fast(ct[j]);
        const float dc = dh * go * (1.f - tc * tc) + dcs[j];
        dO[j] = dh * tc * go * (1.f - go);
        di[j] = dc * gg * gi * (1.f - gi);
        df[j] = dc * cp[j] * gf * (1.f - gf);
        dg[j] = dc * gi * (1.f - gg * gg);
        dcn[j] = dc * gf;
      }
      *reinterpret_cast<float4*>(p.dc + hoff) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
      *reinterpret_cast<float4*>(p.dc + hoff + 4) = make_float4(dcn[4], dcn[5], dcn[6], dcn[7]);
      gp[0] = make_uint4(pack_bf16x2(di[0], di[1]), pack_bf16x2(di[2], di[3]), pack_bf16x2(di[4], di[5]), pack_bf16x2(di[6], di[7]));
      gp[1] = make_uint4(pack_bf16x2(df[0], df[1]), pack_bf16x2(df[2], df[3]), pack_bf16x2(df[4], df[5]), pack_bf16x2(df[6], df[7]));
      gp[2] = make_uint4(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3]), pack_bf16x2(dg[4], dg[5]), pack_bf16x2(dg[6], dg[7]));
      gp[3] = make_uint4(pack_bf16x2(dO[0], dO[1]), pack_bf16x2(dO[2], dO[3]), pack_bf16x2(dO[4], dO[5]), pack_bf16x2(dO[6], dO[7]));
    }
  }
};

// ------------------------------------------------------------------------------------------ projection + L2 norm
// y[b,:] = W h_last[b,:] + bias; emb = y/|y|   (speech_embedder_net.py:31-32; fp32; 8 batch rows per CTA)
constexpr int kProjRows = 8;
__global__ void __launch_bounds__(256) proj_norm_kernel(const float* __restrict__ h, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        float* __restrict__ inv_norm, float* __restrict__ emb, int B,
                                                        int H, int P) {
  extern __shared__ float sm[];
  float* hs = sm;                        // [kProjRows][H]
  float* ys = sm + kProjRows * H;        // [kProjRows][P]
  __shared__ float inv[kProjRows];
  const int b0 = blockIdx.x * kProjRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kProjRows * H; i += 256) {
    const int r = i / H;
    hs[i] = (b0 + r < B) ? h[(size_t)(b0 + r) * H + (i % H)] : 0.f;
  }
  __syncthreads();
  for (int p = warp; p < P; p += 8) {
    float acc[kProjRows];
#pragma unroll
    for (int r = 0; r < kProjRows; ++r) acc[r] = 0.f;
    for (int k = lane; k < H; k += 32) {
      const float wv = W[(size_t)p * H + k];
#pragma unroll
      for (int r = 0; r < kProjRows; ++r) acc[r] += wv * hs[r * H + k];
    }
#pragma unroll
    for (int r = 0; r < kProjRows; ++r) {
      float v = acc[r];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) ys[r * P + p] = v + bias[p];
    }
  }
  __syncthreads();
  if (warp < kProjRows) {
    float ss = 0.f;
    for (int p = lane; p < P; p += 32) ss += ys[warp * P + p] * ys[warp * P + p];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) inv[warp] = 1.0f / sqrtf(ss);          // no epsilon (speech_embedder_net.py:32)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kProjRows * P; i += 256) {
    const int r = i / P, b = b0 + r;
    if (b < B) {
      y[(size_t)b * P + (i % P)] = ys[i];
      emb[(size_t)b * P + (i % P)] = ys[i] * inv[r];
      if ((i % P) == 0) inv_norm[b] = inv[r];
    }
  }
}
// dy = (de - (de.e) e) / |y|  with e = y/|y|
__global__ void norm_bwd_kernel(const float* __restrict__ de, const float* __restrict__ y,
                                const float* __restrict__ inv_norm, float* __restrict__ dy, int B, int P) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float inv = inv_norm[b];
  float dot = 0.f;
  for (int p = lane; p < P; p += 32) dot += de[(size_t)b * P + p] * y[(size_t)b * P + p] * inv;
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int p = lane; p < P; p += 32) {
    const float e = y[(size_t)b * P + p] * inv;
    dy[(size_t)b * P + p] = (de[(size_t)b * P + p] - dot * e) * inv;
  }
}
// Small fp32 SIMT GEMM: C[M,N] = op(A) op(B); 32x32 tiles.  ta: A stored [K,M]; tb: B stored [N,K].
__global__ void __launch_bounds__(256) sgemm_small_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                          float* __restrict__ C, int M, int N, int K, int ta, int tb) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      // As[i][tx] = A(m0+i, k0+tx) ; Bs[i][tx] = B(k0+i, n0+tx)
      int m = m0 + i, k = k0 + tx;
      float av = 0.f;
      if (ta) { int mm = m0 + tx, kk = k0 + i; if (mm < M && kk < K) av = A[(size_t)kk * M + mm]; As[tx][i] = av; }
      else { if (m < M && k < K) av = A[(size_t)m * K + k]; As[i][tx] = av; }
      float bv = 0.f;
      if (tb) { int nn = n0 + i, kk = k0 + tx; if (nn < N && kk < K) bv = Bm[(size_t)nn * K + kk]; Bs[tx][i] = bv; }
      else { int kk = k0 + i, nn = n0 + tx; if (kk < K && nn < N) bv = Bm[(size_t)kk * N + nn]; Bs[i][tx] = bv; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float bv = Bs[k][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += As[ty + 8 * r][k] * bv;
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty + 8 * r, n = n0 + tx;
    if (m < M && n < N) C[(size_t)m * N + n] = acc[r];
  }
}
static void sgemm_small(const float* A, const float* B, float* C, int M, int N, int K, int ta, int tb, cudaStream_t s) {
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  sgemm_small_kernel<<<grid, 256, 0, s>>>(A, B, C, M, N, K, ta, tb);
}
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += x[(size_t)r * cols + c];
  out[c] = acc;
}
// Column sums of dG [rows, 4H] (bf16, packed columns) -> partial sums per row chunk, then unpack + reduce.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part,
                                                          int64_t rows, int cols) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * kColsumRows;
  const int64_t r1 = r0 + kColsumRows < rows ? r0 + kColsumRows : rows;
  float a0 = 0.f, a1 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(x + r * cols + c);
    a0 += bf16_lo_of(v); a1 += bf16_hi_of(v);
  }
  part[(size_t)blockIdx.y * cols + c] = a0;
  part[(size_t)blockIdx.y * cols + c + 1] = a1;
}
__global__ void bias_grad_finish_kernel(const float* __restrict__ part, int chunks, int H, float* __restrict__ g_ih,
                                        float* __restrict__ g_hh) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= 4 * H) return;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += part[(size_t)c * 4 * H + p];
  const int r = ((p & 31) >> 3) * H + (p >> 5) * 8 + (p & 7);
  g_ih[r] = acc;       // b_ih and b_hh enter the pre-activation as a sum: identical gradients
  g_hh[r] = acc;
}

// ------------------------------------------------------------------------------------------ drivers
#define SVB_TRY(expr) do { int _e = (expr); if (_e != SVB_OK) return _e; } while (0)
#define SVB_CUDA(what) do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { set_error(what, _e); return SVB_ERR_CUDA; } } while (0)

static int check_dims(const Dims& d) {
  if (d.B < 1 || d.T < 1 || d.I < 1 || d.L < 1 || d.L > 8 || d.P < 1) { set_error("embedder: bad dims", cudaSuccess); return SVB_ERR_ARG; }
  if (d.H % 128 != 0) { set_error("embedder: hidden size must be a multiple of 128", cudaSuccess); return SVB_ERR_UNSUPPORTED; }
  return SVB_OK;
}

template <class Epi, int BN, bool B_MN>
static int launch_step(GemmOperands& ops, const typename Epi::Params& ep, cudaStream_t s) {
  cudaError_t e = launch_tc_gemm<BN, 4, false, B_MN, Epi>(ops, ep, s);
  if (e != cudaSuccess) { set_error("lstm step launch", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}

}  // namespace svb
using namespace svb;

extern "C" int svb_embedder_sizes(int B, int T, int I, int H, int L, int P, int training, size_t* packed_bytes,
                                  size_t* workspace_bytes) {
  Dims d{B, T, I, H, L, P};
  SVB_TRY(check_dims(d));
  if (packed_bytes) *packed_bytes = layout_packed(nullptr, I, H, L).bytes;
  if (workspace_bytes) *workspace_bytes = layout_work(nullptr, d, training).bytes;
  return SVB_OK;
}

// params: host array of 4L device pointers in nn.LSTM order (weight_ih, weight_hh, bias_ih, bias_hh per layer).
extern "C" int svb_embedder_pack_weights(const float* const* params, void* packed, int I, int H, int L, void* stream) {
  Dims d{1, 1, I, H, L, 1};
  SVB_TRY(check_dims(d));
  PackedW pw = layout_packed(static_cast<char*>(packed), I, H, L);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (int l = 0; l < L; ++l)
    pack_weights_kernel<<<4 * H, 128, 0, s>>>(params[4 * l], params[4 * l + 1], params[4 * l + 2], params[4 * l + 3], pw.l[l], H);
  SVB_CUDA("pack_weights");
  return SVB_OK;
}

// x: (B, T, I) batch-first, x_dtype 0 = float32, 1 = float64.  emb: (B, P) float32.
extern "C" int svb_embedder_forward(const void* x, int x_dtype, const void* packed, const float* proj_w,
                                    const float* proj_b, float* emb, void* workspace, int B, int T, int I, int H, int L,
                                    int P, int training, int rec_terms, void* stream) {
  Dims d{B, T, I, H, L, P};
  SVB_TRY(check_dims(d));
  if (!x || !packed || !proj_w || !proj_b || !emb || !workspace || rec_terms < 1 || rec_terms > 3) return SVB_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  PackedW pw = layout_packed(const_cast<char*>(static_cast<const char*>(packed)), I, H, L);
  Work w = layout_work(static_cast<char*>(workspace), d, training);
  const int Ip0 = round8(I);
  const size_t BH = (size_t)B * H;
  {
    const size_t n = (size_t)T * B * Ip0;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (x_dtype == 0) prep_x_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), w.x_hi, w.x_lo, B, T, I, Ip0);
    else if (x_dtype == 1) prep_x_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(x), w.x_hi, w.x_lo, B, T, I, Ip0);
    else return SVB_ERR_ARG;
    SVB_CUDA("prep_x");
  }
  for (int l = 0; l < L; ++l) {
    cudaMemsetAsync(w.h_hi[l], 0, BH * 2, s);
    cudaMemsetAsync(w.h_lo[l], 0, BH * 2, s);
    cudaMemsetAsync(w.c[l], 0, BH * 4, s);
    if (training) cudaMemsetAsync(w.gates[l] + (size_t)T * B * 4 * H, 0, (size_t)B * 4 * H * 2, s);
  }
  for (int l = 0; l < L; ++l) {
    const LayerW& lw = pw.l[l];
    // ---- input projection over all frames: gin[T*B, 4H] = X W_ih^T + bias (3-term split bf16)
    {
      const __nv_bfloat16* xh = l == 0 ? w.x_hi : w.h_hi[l - 1] + BH;   // slots 1..T of the layer below
      const __nv_bfloat16* xl = l == 0 ? w.x_lo : w.h_lo[l - 1] + BH;
      GemmOperands ops;
      memset(&ops, 0, sizeof(ops));
      ops.nterms = 3; ops.M = T * B; ops.N = 4 * H; ops.K = lw.Ip;
      const void* As[3] = {xh, xl, xh};
      const void* Bs[3] = {lw.wih_hi, lw.wih_hi, lw.wih_lo};
      for (int t = 0; t < 3; ++t) {
        SVB_TRY(make_operand_map(&ops.ta[t], As[t], T * B, lw.Ip, lw.Ip, 0, kBM));
        SVB_TRY(make_operand_map(&ops.tb[t], Bs[t], 4 * H, lw.Ip, lw.Ip, 0, 128));
      }
      EpiStoreF32::Params ep{w.gin, lw.bias, (int64_t)4 * H, 4 * H, 0};
      cudaError_t e = launch_tc_gemm<128, 4, false, false, EpiStoreF32>(ops, ep, s);
      if (e != cudaSuccess) { set_error("input projection", e); return SVB_ERR_CUDA; }
    }
    // ---- recurrence: one fused GEMM + cell kernel per frame
    GemmOperands ops;
    memset(&ops, 0, sizeof(ops));
    // terms: h_hi W_hi (+ h_hi W_lo (+ h_lo W_hi)): rec_terms > 1 buys accuracy for large-magnitude weights
    ops.nterms = rec_terms; ops.M = B; ops.N = 4 * H; ops.K = H;
    SVB_TRY(make_tmap_bf16(&ops.ta[0], w.h_hi[l], H, B, T + 1, H, BH, kBM));
    SVB_TRY(make_operand_map(&ops.tb[0], lw.whh_hi, 4 * H, H, H, 0, 128));
    ops.ta[1] = ops.ta[0];
    SVB_TRY(make_operand_map(&ops.tb[1], lw.whh_lo, 4 * H, H, H, 0, 128));
    SVB_TRY(make_tmap_bf16(&ops.ta[2], w.h_lo[l], H, B, T + 1, H, BH, kBM));
    ops.tb[2] = ops.tb[0];
    for (int t = 0; t < T; ++t) {
      ops.za[0] = ops.za[1] = ops.za[2] = t;
      EpiLstmFwd::Params ep;
      ep.gin = w.gin + (size_t)t * B * 4 * H;
      ep.c_prev = w.c[l] + (size_t)(training ? t : (t & 1)) * BH;
      ep.c_out = w.c[l] + (size_t)(training ? t + 1 : ((t + 1) & 1)) * BH;
      ep.h_hi = w.h_hi[l] + (size_t)(t + 1) * BH;
      ep.h_lo = (l + 1 < L || rec_terms > 2) ? w.h_lo[l] + (size_t)(t + 1) * BH : nullptr;
      ep.gates = training ? w.gates[l] + (size_t)t * B * 4 * H : nullptr;
      ep.h_f32 = (l == L - 1 && t == T - 1) ? w.h_last : nullptr;
      ep.H = H;
      SVB_TRY((launch_step<EpiLstmFwd, 128, false>(ops, ep, s)));
    }
  }
  proj_norm_kernel<<<(B + kProjRows - 1) / kProjRows, 256, (size_t)kProjRows * (H + P) * 4, s>>>(
      w.h_last, proj_w, proj_b, w.y, w.inv_norm, emb, B, H, P);
  SVB_CUDA("proj_norm");
  return SVB_OK;
}

// demb: (B, P).  grads: host array of 4L + 2 device pointers (weight_ih, weight_hh, bias_ih, bias_hh per layer,
// then projection.weight, projection.bias), each fp32 in the parameter's own layout; written, not accumulated.
extern "C" int svb_embedder_backward(const float* demb, const void* packed, const float* proj_w, float* const* grads,
                                     void* workspace, int B, int T, int I, int H, int L, int P, void* stream) {
  Dims d{B, T, I, H, L, P};
  SVB_TRY(check_dims(d));
  if (!demb || !packed || !proj_w || !grads || !workspace) return SVB_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  PackedW pw = layout_packed(const_cast<char*>(static_cast<const char*>(packed)), I, H, L);
  Work w = layout_work(static_cast<char*>(workspace), d, 1);
  const size_t BH = (size_t)B * H;
  const int TB = T * B;
  // ---- projection + norm backward (fp32)
  norm_bwd_kernel<<<(B + 7) / 8, 256, 0, s>>>(demb, w.y, w.inv_norm, w.dy, B, P);
  sgemm_small(w.dy, w.h_last, grads[4 * L], P, H, B, 1, 0, s);          // dW_proj[P,H] = dy^T h_last
  colsum_f32_kernel<<<(P + 127) / 128, 128, 0, s>>>(w.dy, grads[4 * L + 1], B, P);
  sgemm_small(w.dy, proj_w, w.dh_last, B, H, P, 0, 0, s);               // dh_last[B,H] = dy W_proj
  SVB_CUDA("projection backward");
  for (int l = L - 1; l >= 0; --l) {
    const LayerW& lw = pw.l[l];
    cudaMemsetAsync(w.dc, 0, BH * 4, s);
    // ---- BPTT: dh_t(rec) = dG_{t+1} W_hh  fused with the gate backward
    GemmOperands ops;
    memset(&ops, 0, sizeof(ops));
    ops.nterms = 1; ops.M = B; ops.N = H; ops.K = 4 * H;
    SVB_TRY(make_tmap_bf16(&ops.ta[0], w.gates[l], 4 * H, B, T + 1, 4 * H, (size_t)B * 4 * H, kBM));
    SVB_TRY(make_operand_map(&ops.tb[0], lw.whh_hi, H, 4 * H, H, 1, 0));   // [K=4H rows, N=H]: MN-major
    for (int t = T - 1; t >= 0; --t) {
      ops.za[0] = t + 1;
      EpiLstmBwd::Params ep;
      ep.gates = w.gates[l] + (size_t)t * B * 4 * H;
      ep.c_t = w.c[l] + (size_t)(t + 1) * BH;
      ep.c_prev = w.c[l] + (size_t)t * BH;
      ep.dc = w.dc;
      ep.dh_above = (l == L - 1) ? (t == T - 1 ? w.dh_last : nullptr) : w.dh_above + (size_t)t * BH;
      ep.H = H;
      ep.use_acc = 1;
      SVB_TRY((launch_step<EpiLstmBwd, 128, true>(ops, ep, s)));
    }
    // ---- weight gradients: dW[4H, K] = dG^T X over all T*B rows (both operands MN-major), rows unpacked on store
    const __nv_bfloat16* xin = l == 0 ? w.x_hi : w.h_hi[l - 1] + BH;
    {
      GemmOperands g;
      memset(&g, 0, sizeof(g));
      g.nterms = 1; g.M = 4 * H; g.N = H; g.K = TB;
      SVB_TRY(make_operand_map(&g.ta[0], w.gates[l], 4 * H, TB, 4 * H, 1, 0));
      SVB_TRY(make_operand_map(&g.tb[0], w.h_hi[l], H, TB, H, 1, 0));           // h_{t-1}: slots 0..T-1
      EpiStoreF32::Params ep{grads[4 * l + 1], nullptr, (int64_t)H, H, H};
      cudaError_t e = launch_tc_gemm<128, 4, true, true, EpiStoreF32>(g, ep, s);
      if (e != cudaSuccess) { set_error("dW_hh", e); return SVB_ERR_CUDA; }
      g.N = lw.I;
      SVB_TRY(make_operand_map(&g.tb[0], xin, lw.Ip, TB, lw.Ip, 1, 0));
      EpiStoreF32::Params ep2{grads[4 * l], nullptr, (int64_t)lw.I, lw.I, H};
      e = launch_tc_gemm<128, 4, true, true, EpiStoreF32>(g, ep2, s);
      if (e != cudaSuccess) { set_error("dW_ih", e); return SVB_ERR_CUDA; }
    }
    {
      const int chunks = (TB + kColsumRows - 1) / kColsumRows;
      dim3 grid((4 * H / 2 + 255) / 256, chunks);
      colsum_bf16_kernel<<<grid, 256, 0, s>>>(w.gates[l], w.colsum_part, TB, 4 * H);
      bias_grad_finish_kernel<<<(4 * H + 255) / 256, 256, 0, s>>>(w.colsum_part, chunks, H, grads[4 * l + 2], grads[4 * l + 3]);
      SVB_CUDA("bias grads");
    }
    // ---- gradient w.r.t. the layer input = dh_above of the layer below: dX[T*B, H] = dG W_ih
    if (l > 0) {
      GemmOperands g;
      memset(&g, 0, sizeof(g));
      g.nterms = 1; g.M = TB; g.N = H; g.K = 4 * H;
      SVB_TRY(make_operand_map(&g.ta[0], w.gates[l], TB, 4 * H, 4 * H, 0, kBM));
      SVB_TRY(make_operand_map(&g.tb[0], lw.wih_hi, H, 4 * H, lw.Ip, 1, 0));
      EpiStoreF32::Params ep{w.dh_above, nullptr, (int64_t)H, H, 0};
      cudaError_t e = launch_tc_gemm<128, 4, false, true, EpiStoreF32>(g, ep, s);
      if (e != cudaSuccess) { set_error("dX", e); return SVB_ERR_CUDA; }
    }
  }
  return SVB_OK;
}
