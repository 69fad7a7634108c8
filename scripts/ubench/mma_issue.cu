// Microbenchmark: cost of the per-group bookkeeping around 8 back-to-back TS tcgen05.mma (M=128, N=64, K=16).
#include "../../pytorch_speaker_verification_b200/csrc/sm100.cuh"
#include <cstdio>
using namespace svb;

// mode bits: 1 commit to an mbarrier after each group, 2 try_wait on an already-completed barrier before each group,
// 4 tcgen05.fence::after_thread_sync before each group, 8 elect_one + __syncwarp around each group (whole warp loops)
template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, done_bar, ready_bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done_bar, 1); mbar_init(&ready_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x == 0) mbar_arrive(&ready_bar);     // phase 0 of ready_bar completes: parity-0 waits pass immediately
  __syncthreads();
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0, 0);
    const uint64_t d0 = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint32_t lo0 = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
    const bool whole_warp = MODE & 8;
    if (whole_warp || elect_one()) {
      long long t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        if (MODE & 2) mbar_wait(&ready_bar, 0);
        if (MODE & 4) tc_fence_after();
        if (!whole_warp || elect_one()) {
          const uint32_t lo = lo0 + (g & 7) * 1024;
          if (!(MODE & 16))
#pragma unroll
          for (int q = 0; q < 8; ++q)
            umma_f16_ts_lohi(tmem + 384, tmem + (g % 6) * 64 + q * 8, lo + (q >> 2) * 512 + (q & 3) * 2, hi, idesc, 1u);
          if (MODE & 1) umma_commit(&bar);
        }
        if (whole_warp) __syncwarp();
      }
      long long t1 = clock64();
      if (!whole_warp || elect_one()) {
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE>
void run() {
  long long* d; cudaMalloc(&d, 16);
  auto k = bench<MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int groups = 600;
  k<<<148, 128, 160 * 1024>>>(d, groups);
  k<<<148, 128, 160 * 1024>>>(d, groups);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("mode %2d (1 commit, 2 try_wait, 4 fence, 8 whole-warp elect): issue %.1f clk/group, complete %.1f clk/group of 8 MMAs (%s)\n",
         MODE, (double)h[0] / groups, (double)h[1] / groups, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0>(); run<1>(); run<2>(); run<4>(); run<3>(); run<7>(); run<15>(); run<8>();
  printf("-- without MMAs (bit 16): raw bookkeeping cost per group\n");
  run<16 + 1>(); run<16 + 2>(); run<16 + 4>(); run<16 + 3>(); run<16 + 7>(); run<16 + 15>();
  return 0;
}
