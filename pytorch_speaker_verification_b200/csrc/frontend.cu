// Log-mel front end for sm_100a: PCM -> (nmels, T) log10 mel power spectrogram, the
// input format of the hot path.  Replaces data_preprocess.py:41-45 / dvector_create.py:43-47 of the reference
// (librosa.core.stft with center=True reflect padding, periodic Hann window zero-padded to n_fft; |.|^2;
// librosa.filters.mel; log10(. + 1e-6)).  PARITY UNPINNED: librosa is not available in the build container, the
// oracle (oracle/frontend.py) restates its documented algorithm.
//
// One CTA = 8 frames.  The 512-point real DFT is evaluated directly (thread = frequency bin): per sample one complex
// rotation of the thread's running twiddle (re-seeded from an exact table every 32 samples, so the recurrence error
// stays below 1e-6) and 2 FMAs per frame, all frames read as shared-memory broadcasts; 205 k MACs per frame is
// ~0.03 % of what the LSTM spends on the same frame, so an FFT would buy nothing.  Then the 40 x 257 mel projection
// and the log from shared memory.
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace svb {
void set_error(const char* what, cudaError_t e);

constexpr int kFeNfft = 512, kFeBins = kFeNfft / 2 + 1, kFeFrames = 8, kFeThreads = 288, kFeReseed = 32;

__global__ void __launch_bounds__(kFeThreads) logmel_kernel(const float* __restrict__ y, long long n, int n_frames, int hop,
                                                            const float* __restrict__ win, int w0, int w1,
                                                            const float2* __restrict__ tw, const float* __restrict__ melw,
                                                            int nmels, float* __restrict__ out) {
  __shared__ float fr[kFeFrames][kFeNfft];          // windowed frames
  __shared__ float pw[kFeFrames][kFeBins + 3];      // power spectra
  const int t0 = blockIdx.x * kFeFrames;
  for (int i = threadIdx.x; i < kFeFrames * kFeNfft; i += kFeThreads) {
    const int f = i / kFeNfft, m = i % kFeNfft, t = t0 + f;
    float v = 0.f;
    if (t < n_frames && m >= w0 && m < w1) {
      long long idx = (long long)t * hop + m - kFeNfft / 2;       // center=True: frame t is centred on sample t*hop
      if (idx < 0) idx = -idx;                                     // reflect padding (no edge repeat)
      if (idx >= n) idx = 2 * (n - 1) - idx;
      if (idx < 0) idx = 0;
      v = y[idx] * win[m];
    }
    fr[f][m] = v;
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < kFeBins) {
    float re[kFeFrames], im[kFeFrames];
#pragma unroll
    for (int f = 0; f < kFeFrames; ++f) { re[f] = 0.f; im[f] = 0.f; }
    const float2 step = tw[k];                                     // exp(-2 pi i k / 512) as (cos, sin)
    float2 z = make_float2(1.f, 0.f);
    for (int m = w0; m < w1; ++m) {
      if (((m - w0) & (kFeReseed - 1)) == 0) z = tw[(k * m) & (kFeNfft - 1)];     // exact table value
#pragma unroll
      for (int f = 0; f < kFeFrames; ++f) {
        const float v = fr[f][m];
        re[f] = fmaf(v, z.x, re[f]);
        im[f] = fmaf(-v, z.y, im[f]);
      }
      const float zx = z.x * step.x - z.y * step.y;
      z.y = z.x * step.y + z.y * step.x;
      z.x = zx;
    }
#pragma unroll
    for (int f = 0; f < kFeFrames; ++f) pw[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nmels * kFeFrames; o += kFeThreads) {
    const int mel = o / kFeFrames, f = o % kFeFrames, t = t0 + f;
    if (t >= n_frames) continue;
    const float* wrow = melw + (size_t)mel * kFeBins;
    float s = 0.f;
    for (int b = 0; b < kFeBins; ++b) s = fmaf(__ldg(wrow + b), pw[f][b], s);
    out[(size_t)mel * n_frames + t] = log10f(s + 1e-6f);
  }
}
}  // namespace svb
using namespace svb;

extern "C" int svb_logmel(const float* y, int64_t n, int hop, const float* window, int w0, int w1, const float* twiddle,
                          const float* mel_w, int nmels, float* out, int n_frames, void* stream) {
  if (n_frames == 0) return SVB_OK;
  if (!y || !window || !twiddle || !mel_w || !out || n < 1 || hop < 1 || nmels < 1 || n_frames < 0 || w0 < 0 ||
      w1 > kFeNfft || w0 >= w1) {
    set_error("svb_logmel: bad argument (n_fft is fixed at 512)", cudaSuccess);
    return SVB_ERR_ARG;
  }
  logmel_kernel<<<(n_frames + kFeFrames - 1) / kFeFrames, kFeThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y, (long long)n, n_frames, hop, window, w0, w1, reinterpret_cast<const float2*>(twiddle), mel_w, nmels, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_logmel", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
