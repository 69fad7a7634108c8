// Fused GE2E loss forward + backward (fp32) for sm_100a.
//
// Replaces utils.py:27-29 (get_centroids), :40-58 (get_utterance_centroids), :72-115 (get_cossim),
// :126-132 (calc_loss), speech_embedder_net.py:43-49 (GE2ELoss.forward) and the autograd graph behind
// loss.backward() (train_speech_embedder.py:62) of the reference.  Closed form: SURVEY.md section 7.3,
// restated in oracle/ge2e.py::ge2e_fwd_bwd.
//
// One cooperative kernel, four phases separated by grid barriers; nothing of size N*M*N*D is ever
// materialised (the reference's repeat() expansions, utils.py:99-104, are 2 x 41.9 MB at N=64, M=10):
//   A  per speaker : utterance sum s_j, unit centroid c^_j, per-row 1/|e|, 1/|u| and leave-one-out cosine
//   B  per speaker : cos[j,i,:] = e^_ji . c^_k (diag replaced), S = w cos + b, stable log-sum-exp, G, A = wG;
//                    row-local part of dE (sum_k A_off c^_k, diagonal term)
//   C  per (8 centroids x 32 dims): dC_k = (sum_rows A_off e^ - q_k c^_k)/|c_k|
//   D  per speaker : dE += dC_j/M + leave-one-out chain; block 0 reduces loss, dw, db deterministically
// Reductions are warp shuffles + fixed-order shared-memory trees (no float atomics -> run-to-run identical).
#include "../../include/svb200.h"
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace svb {
void set_error(const char* what, cudaError_t e);

constexpr float kCosEps = 1e-8f;    // F.cosine_similarity eps (utils.py:91,105)
constexpr float kCosBias = 1e-6f;   // utils.py:114
constexpr float kLogBias = 1e-6f;   // utils.py:129
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kKC = 8;              // centroids per phase-C tile
constexpr int kDT = 32;             // dims per phase-C tile

struct Ge2eArgs {
  const float* E;        // [N, M, D]
  const float* Cext;     // [Nc, D] foreign centroids or null (then Nc == N, centroids of E)
  const float* w;        // device scalars; null -> cosine only
  const float* b;
  const float* dcos;     // upstream gradient of the cosine matrix [N, M, Nc] (get_cossim backward) or null
  const float* gscale;   // device scalar multiplying all gradients (upstream dL) or null (= 1)
  int N, M, D, Nc;
  int need_grad;
  // outputs (nullable)
  float* cos_out;        // [N, M, Nc]  cos + 1e-6
  float* per_out;        // [N, M]
  float* loss_out;       // scalar
  float* dE;             // [N, M, D]
  float* dCext;          // [Nc, D] gradient w.r.t. foreign centroids
  float* dw;
  float* db;
  // workspace
  float* Ehat;           // [N*M, D]
  float* Chat;           // [Nc, D]
  float* Ssum;           // [N, D]
  float* inv_ne;         // [N*M]
  float* inv_nu;         // [N*M]
  float* cosd;           // [N*M] leave-one-out cosine (no bias)
  float* inv_nc;         // [Nc]
  float* cosm;           // [N*M, Nc] cosine without bias, diagonal replaced
  float* Aoff;           // [N*M, Nc] upstream dL/dcos with the diagonal zeroed
  float* adiag;          // [N*M]
  float* rowstat;        // [3, N*M]: per-row loss, dw, db contributions
  float* dC;             // [Nc, D]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum, result broadcast to every thread.  red must hold kWarps floats.
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) t += red[i];
  return t;
}

// ---------------------------------------------------------------------------------------------- phase A
__device__ void phase_a(const Ge2eArgs& a, float* smem) {
  float* s = smem;              // [D]
  float* red = smem + a.D;      // [kWarps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invM = 1.0f / (float)a.M;
  const float invM1 = 1.0f / (float)(a.M - 1);
  for (int j = blockIdx.x; j < a.N; j += gridDim.x) {
    const float* Ej = a.E + (size_t)j * a.M * a.D;
    float cc = 0.f;
    for (int d = threadIdx.x; d < a.D; d += kThreads) {
      float acc = 0.f;
      for (int m = 0; m < a.M; ++m) acc += Ej[(size_t)m * a.D + d];
      s[d] = acc;
      a.Ssum[(size_t)j * a.D + d] = acc;
      const float c = acc * invM;
      cc += c * c;
    }
    if (!a.Cext) {
      cc = block_sum(cc, red);
      const float inc = 1.0f / fmaxf(sqrtf(cc), kCosEps);
      for (int d = threadIdx.x; d < a.D; d += kThreads) a.Chat[(size_t)j * a.D + d] = s[d] * invM * inc;
      if (threadIdx.x == 0) a.inv_nc[j] = inc;
    }
    __syncthreads();
    for (int m = warp; m < a.M; m += kWarps) {
      const float* e = Ej + (size_t)m * a.D;
      float ee = 0.f, uu = 0.f, eu = 0.f;
      for (int d = lane; d < a.D; d += 32) {
        const float x = e[d];
        const float u = (s[d] - x) * invM1;
        ee += x * x; uu += u * u; eu += x * u;
      }
      ee = warp_sum(ee); uu = warp_sum(uu); eu = warp_sum(eu);
      const float ine = 1.0f / fmaxf(sqrtf(ee), kCosEps);
      const float inu = 1.0f / fmaxf(sqrtf(uu), kCosEps);
      const size_t row = (size_t)j * a.M + m;
      for (int d = lane; d < a.D; d += 32) a.Ehat[row * a.D + d] = e[d] * ine;
      if (lane == 0) {
        a.inv_ne[row] = ine;
        a.inv_nu[row] = inu;
        a.cosd[row] = eu * ine * inu;
      }
    }
    __syncthreads();
  }
  if (a.Cext) {   // foreign centroids (EER: train_speech_embedder.py:127-129): normalise them once
    for (int k = blockIdx.x * kWarps + warp; k < a.Nc; k += gridDim.x * kWarps) {
      const float* c = a.Cext + (size_t)k * a.D;
      float cc = 0.f;
      for (int d = lane; d < a.D; d += 32) cc += c[d] * c[d];
      cc = warp_sum(cc);
      const float inc = 1.0f / fmaxf(sqrtf(cc), kCosEps);
      for (int d = lane; d < a.D; d += 32) a.Chat[(size_t)k * a.D + d] = c[d] * inc;
      if (lane == 0) a.inv_nc[k] = inc;
    }
  }
}

// ---------------------------------------------------------------------------------------------- phase B
__device__ void phase_b(const Ge2eArgs& a, float* smem) {
  const int M = a.M, D = a.D, Nc = a.Nc;
  float* rows = smem;                       // [M, D] unit rows e^
  float* tile = rows + (size_t)M * D;       // [M, Nc] cos, then A_off
  float* rstat = tile + (size_t)M * Nc;     // [M, 4]: r_row, a_diag, (unused)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float w = a.w ? *a.w : 0.f, b = a.b ? *a.b : 0.f;
  const float invM1 = 1.0f / (float)(M - 1);
  for (int j = blockIdx.x; j < a.N; j += gridDim.x) {
    const size_t row0 = (size_t)j * M;
    for (int i = threadIdx.x; i < M * D; i += kThreads) rows[i] = a.Ehat[row0 * D + i];
    __syncthreads();
    // cos[m, k] = e^_m . c^_k : one warp per centroid, lanes over d, rows from shared memory
    for (int k = warp; k < Nc; k += kWarps) {
      const float* c = a.Chat + (size_t)k * D;
      for (int m0 = 0; m0 < M; m0 += 8) {
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        for (int d = lane; d < D; d += 32) {
          const float cv = c[d];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (m0 + r < M) acc[r] += cv * rows[(size_t)(m0 + r) * D + d];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float v = warp_sum(acc[r]);
          if (lane == 0 && m0 + r < M) tile[(size_t)(m0 + r) * Nc + k] = v;
        }
      }
    }
    __syncthreads();
    // per-row softmax contrast: one warp per row
    for (int m = warp; m < M; m += kWarps) {
      const size_t row = row0 + m;
      float* t = tile + (size_t)m * Nc;
      const float cd = a.cosd[row];
      if (lane == 0 && j < Nc) t[j] = cd;          // diagonal overwrite (utils.py:113)
      __syncwarp();
      for (int k = lane; k < Nc; k += 32) {
        const float c0 = t[k];
        a.cosm[row * Nc + k] = c0;
        if (a.cos_out) a.cos_out[row * Nc + k] = c0 + kCosBias;
      }
      if (a.w) {
        float mx = -INFINITY;
        for (int k = lane; k < Nc; k += 32) mx = fmaxf(mx, w * (t[k] + kCosBias) + b);
        mx = warp_max(mx);
        float se = 0.f;
        for (int k = lane; k < Nc; k += 32) se += expf(w * (t[k] + kCosBias) + b - mx);
        se = warp_sum(se);
        // log(sum exp S + 1e-6) = mx + log(se + 1e-6 e^{-mx})  (reference has no max-subtraction: same value)
        const float tiny = kLogBias * expf(-mx);
        const float den = se + tiny;
        const float sdiag = w * (cd + kCosBias) + b;
        const float per = -sdiag + mx + logf(den);
        const float inv_den = 1.0f / den;
        float dwp = 0.f, rr = 0.f, ad = 0.f;
        for (int k = lane; k < Nc; k += 32) {
          const float c0 = t[k];
          float g = expf(w * (c0 + kCosBias) + b - mx) * inv_den;
          if (k == j) g -= 1.0f;
          dwp += g * (c0 + kCosBias);
          const float A = w * g;
          float ao = A;
          if (k == j) { ad = A; ao = 0.f; }
          rr += ao * c0;
          t[k] = ao;
          a.Aoff[row * Nc + k] = ao;
        }
        dwp = warp_sum(dwp); rr = warp_sum(rr); ad = warp_sum(ad);
        if (lane == 0) {
          const int NM = a.N * M;
          a.rowstat[row] = per;
          a.rowstat[NM + row] = dwp;
          a.rowstat[2 * NM + row] = -tiny * inv_den;     // sum_k G = -1e-6/den exactly
          if (a.per_out) a.per_out[row] = per;
          a.adiag[row] = ad;
          rstat[m * 4 + 0] = rr;
          rstat[m * 4 + 1] = ad;
        }
      } else if (a.dcos) {   // get_cossim backward: upstream gradient given
        float rr = 0.f, ad = 0.f;
        for (int k = lane; k < Nc; k += 32) {
          const float c0 = t[k];
          const float A = a.dcos[row * Nc + k];
          float ao = A;
          if (k == j) { ad = A; ao = 0.f; }
          rr += ao * c0;
          t[k] = ao;
          a.Aoff[row * Nc + k] = ao;
        }
        rr = warp_sum(rr); ad = warp_sum(ad);
        if (lane == 0) {
          a.adiag[row] = ad;
          rstat[m * 4 + 0] = rr;
          rstat[m * 4 + 1] = ad;
        }
      }
    }
    __syncthreads();
    if (a.need_grad) {
      // row-local gradient: thread per dim, 8 rows of accumulators
      const float* sj = a.Ssum + (size_t)j * D;
      for (int d = threadIdx.x; d < D; d += kThreads) {
        for (int m0 = 0; m0 < M; m0 += 8) {
          float acc[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[r] = 0.f;
          for (int k = 0; k < Nc; ++k) {
            const float cv = a.Chat[(size_t)k * D + d];
#pragma unroll
            for (int r = 0; r < 8; ++r)
              if (m0 + r < M) acc[r] += tile[(size_t)(m0 + r) * Nc + k] * cv;
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int m = m0 + r;
            if (m < M) {
              const size_t row = row0 + m;
              const float eh = rows[(size_t)m * D + d];
              const float ine = a.inv_ne[row];
              const float uh = (sj[d] - a.E[row * D + d]) * invM1 * a.inv_nu[row];
              const float cd = a.cosd[row];
              a.dE[row * D + d] = (acc[r] - rstat[m * 4 + 0] * eh) * ine + rstat[m * 4 + 1] * (uh - cd * eh) * ine;
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- phase C
__device__ void phase_c(const Ge2eArgs& a, float* smem) {
  float* red = smem;                       // [kWarps][kKC][kDT]
  float* qred = red + kWarps * kKC * kDT;  // [kWarps][kKC]
  const int D = a.D, Nc = a.Nc, NM = a.N * a.M;
  const int ktiles = (Nc + kKC - 1) / kKC, dtiles = (D + kDT - 1) / kDT;
  const int rg = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp = row group, lane = dim
  for (int tix = blockIdx.x; tix < ktiles * dtiles; tix += gridDim.x) {
    const int k0 = (tix / dtiles) * kKC, d = (tix % dtiles) * kDT + lane;
    float acc[kKC];
#pragma unroll
    for (int i = 0; i < kKC; ++i) acc[i] = 0.f;
    float q = 0.f;
    for (int row = rg; row < NM; row += kWarps) {
      const float eh = d < D ? a.Ehat[(size_t)row * D + d] : 0.f;
      const float* ao = a.Aoff + (size_t)row * Nc + k0;
#pragma unroll
      for (int i = 0; i < kKC; ++i)
        if (k0 + i < Nc) acc[i] += ao[i] * eh;
      if (lane < kKC && k0 + lane < Nc) q += ao[lane] * a.cosm[(size_t)row * Nc + k0 + lane];
    }
#pragma unroll
    for (int i = 0; i < kKC; ++i) red[(rg * kKC + i) * kDT + lane] = acc[i];
    if (lane < kKC) qred[rg * kKC + lane] = q;
    __syncthreads();
    {
      const int i = threadIdx.x >> 5;      // kWarps == kKC: warp i finishes centroid k0 + i
      const int k = k0 + i;
      if (k < Nc && d < D) {
        float p = 0.f, qq = 0.f;
#pragma unroll
        for (int g = 0; g < kWarps; ++g) { p += red[(g * kKC + i) * kDT + lane]; qq += qred[g * kKC + i]; }
        a.dC[(size_t)k * D + d] = (p - qq * a.Chat[(size_t)k * D + d]) * a.inv_nc[k];
      }
    }
    __syncthreads();
  }
}
static_assert(kWarps == kKC, "phase C maps one warp per centroid of the tile");

// ---------------------------------------------------------------------------------------------- phase D
__device__ void phase_d(const Ge2eArgs& a, float* smem) {
  float* red = smem;
  const int M = a.M, D = a.D, NM = a.N * a.M;
  const float gs = a.gscale ? *a.gscale : 1.0f;
  if (a.need_grad) {
    const float invM = 1.0f / (float)M, invM1 = 1.0f / (float)(M - 1);
    for (int j = blockIdx.x; j < a.N; j += gridDim.x) {
      const size_t row0 = (size_t)j * M;
      for (int d = threadIdx.x; d < D; d += kThreads) {
        const float sj = a.Ssum[(size_t)j * D + d];
        float sumdu = 0.f;
        for (int m = 0; m < M; ++m) {
          const size_t row = row0 + m;
          const float eh = a.Ehat[row * D + d];
          const float inu = a.inv_nu[row];
          const float uh = (sj - a.E[row * D + d]) * invM1 * inu;
          sumdu += a.adiag[row] * (eh - a.cosd[row] * uh) * inu;
        }
        const float dc = (a.Cext == nullptr) ? a.dC[(size_t)j * D + d] * invM : 0.f;
        for (int m = 0; m < M; ++m) {
          const size_t row = row0 + m;
          const float eh = a.Ehat[row * D + d];
          const float inu = a.inv_nu[row];
          const float uh = (sj - a.E[row * D + d]) * invM1 * inu;
          const float du = a.adiag[row] * (eh - a.cosd[row] * uh) * inu;
          a.dE[row * D + d] = gs * (a.dE[row * D + d] + dc + (sumdu - du) * invM1);
        }
      }
    }
    if (a.Cext && a.dCext) {
      for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < (size_t)a.Nc * D; i += (size_t)gridDim.x * kThreads)
        a.dCext[i] = gs * a.dC[i];
    }
  }
  if (blockIdx.x == 0 && a.w) {
    float l = 0.f, dw = 0.f, db = 0.f;
    for (int r = threadIdx.x; r < NM; r += kThreads) {
      l += a.rowstat[r]; dw += a.rowstat[NM + r]; db += a.rowstat[2 * NM + r];
    }
    l = block_sum(l, red); dw = block_sum(dw, red); db = block_sum(db, red);
    if (threadIdx.x == 0) {
      if (a.loss_out) *a.loss_out = l;
      if (a.dw) *a.dw = gs * dw;
      if (a.db) *a.db = gs * db;
    }
  }
}

__global__ void __launch_bounds__(kThreads) ge2e_fused_kernel(const Ge2eArgs a) {
  extern __shared__ float smem[];
  cg::grid_group grid = cg::this_grid();
  phase_a(a, smem);
  grid.sync();
  phase_b(a, smem);
  if (a.need_grad) {
    grid.sync();
    phase_c(a, smem);
  }
  grid.sync();
  phase_d(a, smem);
}
__global__ void __launch_bounds__(kThreads) ge2e_phase_kernel(const Ge2eArgs a, int phase) {
  extern __shared__ float smem[];
  if (phase == 0) phase_a(a, smem);
  else if (phase == 1) phase_b(a, smem);
  else if (phase == 2) phase_c(a, smem);
  else phase_d(a, smem);
}

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

static size_t carve(Ge2eArgs& a, char* base) {
  const size_t NM = (size_t)a.N * a.M, D = a.D, Nc = a.Nc;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* p = base ? reinterpret_cast<float*>(base + off) : nullptr; off += align_up(nfloats * 4); return p; };
  a.Ehat = take(NM * D); a.Chat = take(Nc * D); a.Ssum = take((size_t)a.N * D);
  a.inv_ne = take(NM); a.inv_nu = take(NM); a.cosd = take(NM); a.inv_nc = take(Nc);
  a.cosm = take(NM * Nc); a.Aoff = take(NM * Nc); a.adiag = take(NM); a.rowstat = take(3 * NM); a.dC = take(Nc * D);
  return off;
}

static size_t smem_bytes(const Ge2eArgs& a) {
  size_t pa = (size_t)a.D + kWarps;
  size_t pb = (size_t)a.M * a.D + (size_t)a.M * a.Nc + (size_t)a.M * 4;
  size_t pc = (size_t)kWarps * kKC * kDT + kWarps * kKC;
  size_t m = pa > pb ? pa : pb;
  m = m > pc ? m : pc;
  return m * sizeof(float);
}

}  // namespace svb

using namespace svb;

extern "C" int svb_ge2e_workspace_bytes(int N, int M, int D, int Nc, size_t* bytes) {
  if (N < 1 || M < 1 || D < 1 || Nc < 1 || !bytes) return SVB_ERR_ARG;
  Ge2eArgs a{};
  a.N = N; a.M = M; a.D = D; a.Nc = Nc;
  *bytes = carve(a, nullptr);
  return SVB_OK;
}

extern "C" int svb_ge2e(const float* E, const float* Cext, int N, int M, int D, int Nc, const float* w,
                        const float* b, const float* dcos, const float* gscale, float* cos_out, float* per_out,
                        float* loss_out, float* dE, float* dCext, float* dw, float* db, void* workspace,
                        size_t workspace_bytes, int fused, void* stream) {
  if (!E || N < 1 || M < 2 || D < 1 || !workspace) { set_error("svb_ge2e: bad argument (M must be >= 2)", cudaSuccess); return SVB_ERR_ARG; }
  if (!Cext && Nc != N) { set_error("svb_ge2e: Nc must equal N without foreign centroids", cudaSuccess); return SVB_ERR_ARG; }
  if ((w == nullptr) != (b == nullptr)) return SVB_ERR_ARG;
  Ge2eArgs a{};
  a.E = E; a.Cext = Cext; a.w = w; a.b = b; a.dcos = dcos; a.gscale = gscale;
  a.N = N; a.M = M; a.D = D; a.Nc = Nc;
  a.need_grad = (dE != nullptr) ? 1 : 0;
  if (a.need_grad && !w && !dcos) { set_error("svb_ge2e: gradient requested without w/b or dcos", cudaSuccess); return SVB_ERR_ARG; }
  a.cos_out = cos_out; a.per_out = per_out; a.loss_out = loss_out; a.dE = dE; a.dCext = dCext; a.dw = dw; a.db = db;
  if (carve(a, static_cast<char*>(workspace)) > workspace_bytes) { set_error("svb_ge2e: workspace too small", cudaSuccess); return SVB_ERR_ARG; }
  const size_t smem = smem_bytes(a);
  if (smem > 200 * 1024) { set_error("svb_ge2e: M*(D+Nc) too large for one CTA's shared memory", cudaSuccess); return SVB_ERR_UNSUPPORTED; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static int max_smem_set = 0, num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if ((int)smem > max_smem_set) {
    cudaError_t e = cudaFuncSetAttribute(ge2e_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ge2e_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("svb_ge2e: cudaFuncSetAttribute", e); return SVB_ERR_CUDA; }
    max_smem_set = (int)smem;
  }
  const int ctiles = ((Nc + kKC - 1) / kKC) * ((D + kDT - 1) / kDT);
  int want = N > ctiles ? N : ctiles;
  if (fused) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ge2e_fused_kernel, kThreads, smem);
    if (per_sm < 1) { set_error("svb_ge2e: kernel does not fit", cudaSuccess); return SVB_ERR_UNSUPPORTED; }
    int grid = want < per_sm * num_sms ? want : per_sm * num_sms;
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)ge2e_fused_kernel, dim3(grid), dim3(kThreads), params, smem, s);
    if (e != cudaSuccess) { set_error("svb_ge2e: cooperative launch", e); return SVB_ERR_CUDA; }
  } else {
    ge2e_phase_kernel<<<N, kThreads, smem, s>>>(a, 0);
    ge2e_phase_kernel<<<N, kThreads, smem, s>>>(a, 1);
    if (a.need_grad) ge2e_phase_kernel<<<ctiles, kThreads, smem, s>>>(a, 2);
    ge2e_phase_kernel<<<N, kThreads, smem, s>>>(a, 3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("svb_ge2e: launch", e); return SVB_ERR_CUDA; }
  }
  return SVB_OK;
}

// get_centroids (utils.py:27-29): C[j, d] = mean_m E[j, m, d]; backward: dE[j, m, d] = dC[j, d] / M.
__global__ void centroid_kernel(const float* __restrict__ E, float* __restrict__ C, int N, int M, int D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * D) return;
  const int j = i / D, d = i % D;
  float acc = 0.f;
  for (int m = 0; m < M; ++m) acc += E[((size_t)j * M + m) * D + d];
  C[i] = acc / (float)M;
}
__global__ void centroid_bwd_kernel(const float* __restrict__ dC, float* __restrict__ dE, int N, int M, int D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * M * D) return;
  const int d = i % D;
  const int j = i / ((size_t)M * D);
  dE[i] = dC[(size_t)j * D + d] / (float)M;
}
extern "C" int svb_centroids(const float* E, float* C, int N, int M, int D, void* stream) {
  if (!E || !C || N < 1 || M < 1 || D < 1) return SVB_ERR_ARG;
  const size_t n = (size_t)N * D;
  centroid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(E, C, N, M, D);
  return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}
extern "C" int svb_centroids_bwd(const float* dC, float* dE, int N, int M, int D, void* stream) {
  if (!dC || !dE || N < 1 || M < 1 || D < 1) return SVB_ERR_ARG;
  const size_t n = (size_t)N * M * D;
  centroid_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dC, dE, N, M, D);
  return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

// calc_loss (utils.py:126-132) on a caller-supplied similarity matrix S [N, M, Nc]; optional gradient
// dS = gscale * (softmax-with-bias - onehot).  One warp per row.
__global__ void calc_loss_kernel(const float* __restrict__ S, float* __restrict__ per, float* __restrict__ dS,
                                 const float* __restrict__ gscale, int N, int M, int Nc) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N * M) return;
  const int j = row / M;
  const float* s = S + (size_t)row * Nc;
  float mx = -INFINITY;
  for (int k = lane; k < Nc; k += 32) mx = fmaxf(mx, s[k]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int k = lane; k < Nc; k += 32) se += expf(s[k] - mx);
  se = warp_sum(se);
  const float den = se + kLogBias * expf(-mx);
  if (lane == 0) per[row] = -s[j] + mx + logf(den);
  if (dS) {
    const float gs = gscale ? *gscale : 1.0f;
    for (int k = lane; k < Nc; k += 32) dS[(size_t)row * Nc + k] = gs * (expf(s[k] - mx) / den - (k == j ? 1.f : 0.f));
  }
}
__global__ void sum_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
  __shared__ float red[8];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) v += x[i];
  v = block_sum(v, red);
  if (threadIdx.x == 0) *out = v;
}
extern "C" int svb_calc_loss(const float* S, int N, int M, int Nc, float* per_out, float* loss_out, float* dS,
                             const float* gscale, void* stream) {
  if (!S || !per_out || N < 1 || M < 1 || Nc < N) return SVB_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  calc_loss_kernel<<<(N * M + 7) / 8, 256, 0, st>>>(S, per_out, dS, gscale, N, M, Nc);
  if (loss_out) sum_kernel<<<1, 256, 0, st>>>(per_out, loss_out, N * M);
  return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}
