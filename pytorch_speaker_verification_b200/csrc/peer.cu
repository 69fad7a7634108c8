// Exchange steps of the multi-GPU GE2E over NVLink PEER MEMORY (one process per GPU; the buffers are symmetric
// allocations whose peer addresses every rank holds, torch.distributed._symmetric_memory): the ranks read each
// other's centroids and centroid-gradient partials with plain loads instead of going through NCCL (an all-gather of
// 65 KB and an all-reduce of 524 KB per step at 8 GPUs: two collectives whose cost is latency, not bandwidth).
// New functionality: the reference (speech_embedder_net.py:43-49) is single-device.  Both kernels add in RANK ORDER:
// every rank computes bit-identical sums, run to run.  The same two kernels make a two-shot ALL-REDUCE of the 48.5 MB of
// parameter gradients (reduce-scatter: every rank sums its slice over all peers in place; all-gather: every rank reads
// the reduced slices back), see dist.PeerExchange.allreduce_.
#include "../../include/svb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace svb {
void set_error(const char* what, cudaError_t e);

// out[r * n + i] = peer_r[off + r * rank_stride + i]   (n % 4 == 0, 16-byte aligned)
__global__ void __launch_bounds__(256) peer_gather_kernel(const float* const* __restrict__ peers, int world, size_t off,
                                                          size_t rank_stride, size_t n, float* __restrict__ out) {
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4 * world; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4);
    const size_t k = i - (size_t)r * n4;
    const float4 v = *(reinterpret_cast<const float4*>(peers[r] + off + (size_t)r * rank_stride) + k);
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// seg[i] = sum_r peer_r[off + seg_off + i] (i < seg_n, seg_n % 4 == 0); tail[k] = sum_r peer_r[off + tail_off + k]
__global__ void __launch_bounds__(256) peer_reduce_kernel(const float* const* __restrict__ peers, int world, size_t off,
                                                          size_t seg_off, size_t seg_n, size_t tail_off, int tail_n,
                                                          float* __restrict__ seg, float* __restrict__ tail) {
  const size_t n4 = seg_n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = *(reinterpret_cast<const float4*>(peers[r] + off + seg_off) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(seg)[i] = acc;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail_n) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r) acc += peers[r][off + tail_off + threadIdx.x];
    tail[threadIdx.x] = acc;
  }
}

}  // namespace svb
using namespace svb;

extern "C" int svb_peer_gather(const void* peer_ptrs_dev, int world, size_t offset_floats, size_t rank_stride_floats,
                               size_t n_floats, float* out, void* stream) {
  if (!peer_ptrs_dev || !out || world < 1 || (n_floats & 3) || (offset_floats & 3) || (rank_stride_floats & 3)) {
    set_error("svb_peer_gather: bad argument (lengths and offsets in multiples of 4 floats)", cudaSuccess);
    return SVB_ERR_ARG;
  }
  const size_t n4 = n_floats / 4 * world;
  const unsigned grid = (unsigned)((n4 + 255) / 256 < 592 ? (n4 + 255) / 256 : 592);
  peer_gather_kernel<<<grid ? grid : 1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const float* const*>(peer_ptrs_dev), world, offset_floats, rank_stride_floats, n_floats, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_peer_gather", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}

extern "C" int svb_peer_reduce(const void* peer_ptrs_dev, int world, size_t offset_floats, size_t seg_offset,
                               size_t seg_floats, size_t tail_offset, int tail_floats, float* seg_out, float* tail_out,
                               void* stream) {
  if (!peer_ptrs_dev || !seg_out || world < 1 || (seg_floats & 3) || ((offset_floats + seg_offset) & 3) || tail_floats < 0 ||
      tail_floats > 256 || (tail_floats && !tail_out)) {
    set_error("svb_peer_reduce: bad argument", cudaSuccess);
    return SVB_ERR_ARG;
  }
  const size_t n4 = seg_floats / 4;
  const unsigned grid = (unsigned)((n4 + 255) / 256 < 592 ? (n4 + 255) / 256 : 592);
  peer_reduce_kernel<<<grid ? grid : 1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const float* const*>(peer_ptrs_dev), world, offset_floats, seg_offset, seg_floats, tail_offset,
      tail_floats, seg_out, tail_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("svb_peer_reduce", e); return SVB_ERR_CUDA; }
  return SVB_OK;
}
