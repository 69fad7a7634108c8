// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) for A from TMEM (TS) vs shared memory (SS)
// and N = 64 / 128 / 256, issued back to back by one thread.   nvcc -gencode arch=compute_100a,code=sm_100a
#include "../../pytorch_speaker_verification_b200/csrc/sm100.cuh"
#include <cstdio>
using namespace svb;

template <int N, bool TS, int NCTA>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_bf16_ts(tmem + 256, tmem + (i & 7) * 32 + k * 8, umma_desc_kmajor_sw128(sb + k * 32), idesc, 1u);
        else umma_bf16_ss(tmem + 256, umma_desc_kmajor_sw128(sa + k * 32), umma_desc_kmajor_sw128(sb + k * 32), idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N, bool TS>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  auto k = bench<N, TS, 1>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  k<<<grid, 128, 64 * 1024>>>(d, iters);
  k<<<grid, 128, 64 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-10s N=%3d grid=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA (%s)\n", name, N, grid, (double)h[0] / (4.0 * iters),
         (double)h[1] / (4.0 * iters), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int grid : {1, 148}) {
    run<64, true>("TS", grid); run<128, true>("TS", grid); run<256, true>("TS", grid);
    run<64, false>("SS", grid); run<128, false>("SS", grid); run<256, false>("SS", grid);
  }
  return 0;
}
