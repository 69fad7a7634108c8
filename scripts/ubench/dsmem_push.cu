// Microbenchmark: all-to-all exchange of 4 KB messages inside a 4-CTA cluster (the split-K partial exchange of
// wbptt.cuh: every CTA sends one quarter to each of its three peers per tile), cycles per round for
//   A  st.shared::cluster.v4 from 6 warps (2 per destination, 2 x 2 16-byte stores per lane) + remote mbarrier arrive
//   B  cp.async.bulk.shared::cluster.shared::cta (one thread, 3 copies of 4 KB, complete_tx on the receiver's mbarrier)
//   C  as B with 6 copies of 2 KB
// All 30 clusters of the BPTT kernel's grid run at once (120 CTAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_push dsmem_push.cu
#include "../../pytorch_speaker_verification_b200/csrc/sm100.cuh"
#include <cstdio>
using namespace svb;

__device__ __forceinline__ void st_cluster_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_cta(smem_u32(bar), cta)) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster_addr, uint32_t src, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(long long* out, int iters) {
  __shared__ __align__(128) uint8_t recv[2][3 * 4096];
  __shared__ __align__(128) uint8_t send[1][3 * 4096];
  __shared__ uint64_t full[2], freeb[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&full[b], MODE == 0 ? 6 : 1); mbar_init(&freeb[b], 3); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 3 * 4096 / 4; i += 256) reinterpret_cast<uint32_t*>(send)[i] = i;
  fence_proxy_async_smem();
  cluster_sync_all();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    const uint32_t par = (it >> 1) & 1;
    if (MODE == 0) {
      if (warp < 6) {
        const int dsti = warp >> 1, half = warp & 1;          // destination (s + 1 + dsti) & 3, slot of the receiver = 2 - dsti
        const uint32_t q = (s + 1 + dsti) & 3;
        if (it >= 2) while (!try_wait_cluster(&freeb[buf], par ^ 1)) {}
        const uint32_t dst = map_to_cta(smem_u32(&recv[buf][0]) + ((s - q - 1) & 3) * 4096 + half * 2048 + lane * 16, q);
#pragma unroll
        for (int k = 0; k < 4; ++k) st_cluster_u4(dst + k * 512, make_uint4(it, lane, k, warp));
        __syncwarp();
        if (lane == 0) arrive_remote(&full[buf], q);
      }
    } else {
      if (warp == 0 && lane == 0) {
        if (it >= 2) while (!try_wait_cluster(&freeb[buf], par ^ 1)) {}
        for (int dsti = 0; dsti < 3; ++dsti) {
          const uint32_t q = (s + 1 + dsti) & 3;
          const uint32_t slot = (s - q - 1) & 3;
          const uint32_t bar = map_to_cta(smem_u32(&full[buf]), q);
          if (MODE == 1) {
            bulk_push(map_to_cta(smem_u32(&recv[buf][0]) + slot * 4096, q), smem_u32(&send[0][0]) + dsti * 4096, 4096, bar);
          } else {
            bulk_push(map_to_cta(smem_u32(&recv[buf][0]) + slot * 4096, q), smem_u32(&send[0][0]) + dsti * 4096, 2048, bar);
            bulk_push(map_to_cta(smem_u32(&recv[buf][0]) + slot * 4096 + 2048, q), smem_u32(&send[0][0]) + dsti * 4096 + 2048, 2048, bar);
          }
        }
      }
    }
    // receiver: wait for the three quarters, "consume", tell the senders
    if (warp == 7) {
      if (MODE != 0 && lane == 0) mbar_expect_tx(&full[buf], 3 * 4096);     // (arrival count 1 + 12 KB of transactions)
      __syncwarp();
      while (!try_wait_cluster(&full[buf], par)) {}
      uint32_t acc = reinterpret_cast<uint32_t*>(&recv[buf][0])[lane];
      if (acc == 0xdeadbeefu) out[1] = 1;
      __syncwarp();
      if (lane < 3) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_cta(smem_u32(&freeb[buf]), (s + 1 + lane) & 3)) : "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  cluster_sync_all();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

template <int MODE>
static void run(const char* name, long long* d_out, int iters) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(120); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench<MODE>, d_out, iters);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-60s %6lld cycles per round (12 KB out + 12 KB in per CTA)\n", name, h[0]);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  cudaMemset(d_out, 0, 64);
  run<0>("A  st.shared::cluster.v4, 6 warps, remote arrives", d_out, 2000);
  run<1>("B  cp.async.bulk smem->peer smem, 3 x 4 KB, complete_tx", d_out, 2000);
  run<2>("C  cp.async.bulk smem->peer smem, 6 x 2 KB, complete_tx", d_out, 2000);
  return 0;
}
