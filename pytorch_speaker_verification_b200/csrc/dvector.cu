// d-vector extraction helpers (dvector_create.py of the reference) for sm_100a.
//   svb_dvector_windows : :48-52 + :98-99  sliding 24-frame windows (hop 12) of a (nmels, T) log-mel matrix,
//                          emitted as (W, win, nmels) -- a transposing gather, HBM-bound.
//   svb_segment_mean    : :55-73 align_embeddings: per-partition float32 running sum in row order, float32
//                          divide by the count (np.average on float32 rows), widened to float64 on store.
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace svb {
void set_error(const char* what, cudaError_t e);

// S: [nmels, Ttot] (row pitch ldS); win_start[w] = first frame of window w (host-computed from the strict
// "j + win < T" rule per utterance).  out: [W, win, nmels].
__global__ void __launch_bounds__(256) windows_kernel(const float* __restrict__ S, int64_t ldS, int nmels,
                                                      const int* __restrict__ win_start, int W, int win,
                                                      float* __restrict__ out) {
  extern __shared__ float tile[];          // [nmels][win + 1]
  for (int w = blockIdx.x; w < W; w += gridDim.x) {
    const int f0 = win_start[w];
    for (int i = threadIdx.x; i < nmels * win; i += blockDim.x) {
      const int mel = i / win, f = i % win;               // coalesced along frames
      tile[mel * (win + 1) + f] = S[(int64_t)mel * ldS + f0 + f];
    }
    __syncthreads();
    float* o = out + (size_t)w * win * nmels;
    for (int i = threadIdx.x; i < nmels * win; i += blockDim.x) {
      const int f = i / nmels, mel = i % nmels;           // coalesced along mels
      o[i] = tile[mel * (win + 1) + f];
    }
    __syncthreads();
  }
}

__global__ void segment_mean_kernel(const float* __restrict__ emb, int D, const int* __restrict__ seg_off, int P,
                                    double* __restrict__ out) {
  const int p = blockIdx.x;
  const int s = seg_off[p], e = seg_off[p + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int r = s; r < e; ++r) acc = __fadd_rn(acc, emb[(size_t)r * D + d]);
    out[(size_t)p * D + d] = (double)__fdiv_rn(acc, (float)(e - s));
  }
}
}  // namespace svb
using namespace svb;

extern "C" int svb_dvector_windows(const float* S, int64_t ldS, int nmels, const int* win_start, int W, int win,
                                   float* out, void* stream) {
  if (W == 0) return SVB_OK;
  if (!S || !win_start || !out || nmels < 1 || win < 1 || W < 0) return SVB_ERR_ARG;
  const int grid = W < 148 * 8 ? W : 148 * 8;
  windows_kernel<<<grid, 256, (size_t)nmels * (win + 1) * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      S, ldS, nmels, win_start, W, win, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_dvector_windows", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}

extern "C" int svb_segment_mean(const float* emb, int D, const int* seg_offsets, int P, double* out, void* stream) {
  if (P == 0) return SVB_OK;
  if (!emb || !seg_offsets || !out || D < 1 || P < 0) return SVB_ERR_ARG;
  segment_mean_kernel<<<P, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(emb, D, seg_offsets, P, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_segment_mean", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
