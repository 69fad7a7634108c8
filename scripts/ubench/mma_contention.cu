// Microbenchmark: TS tcgen05.mma (M=128, N=64, K=16, fp16) rate while other agents use the SM:
//   bit 1: a warp streams 8 KB bulk copies global -> shared into the operand ring (like the TMA producer)
//   bit 2: 16 warps run MUFU/FMA-heavy code (like the epilogue)
//   bit 4: 8 warps do tcgen05.ld of the other accumulator in a loop
//   bit 8: 16 warps do shared-memory stores/loads (like the staging writes)
#include "../../pytorch_speaker_verification_b200/csrc/sm100.cuh"
#include <cstdio>
using namespace svb;

template <int MODE>
__global__ void __launch_bounds__(640, 1) bench(long long* out, const uint8_t* gsrc, float* sink, int groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar, ld_bar[16];
  __shared__ uint32_t holder;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); for (int i = 0; i < 16; ++i) mbar_init(&ld_bar[i], 1); fence_mbar_init(); stop = 0; }
  if (warp == 17) tmem_alloc<512>(&holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 17) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0, 0);
      const uint64_t d0 = umma_desc_kmajor_sw128(smem_u32(smem));
      const uint32_t lo0 = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
      long long t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const uint32_t lo = lo0 + (g % 3) * 3072;      // 3 stages of 48 KB
#pragma unroll
        for (int q = 0; q < 24; ++q)
          umma_f16_ts_lohi(tmem + 384 + (g & 1) * 64, tmem + (g & 1) * 192 + q * 8, lo + (q >> 2) * 512 + (q & 3) * 2, hi, idesc, 1u);
      }
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 16 && (MODE & 1)) {
    if (elect_one()) {
      int i = 0; uint32_t ph[16] = {0};
      while (!stop) {
        const int b = i & 15;
        if (i >= 16) { mbar_wait(&ld_bar[b], ph[b]); ph[b] ^= 1; }
        mbar_expect_tx(&ld_bar[b], 8192);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + 147456 - 8192 * 16 + b * 8192)), "l"(gsrc + (size_t)((i * 97 + blockIdx.x * 13) % 4096) * 8192), "n"(8192),
                       "r"(smem_u32(&ld_bar[b])) : "memory");
        ++i;
      }
      for (int b = 0; b < 16 && b < i; ++b) mbar_wait(&ld_bar[b], ph[b]);
    }
  } else if (warp < 16) {
    float acc = threadIdx.x * 0.001f;
    float v[16];
    while (!stop) {
      if (MODE & 2) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(0.5f, tanh_approx(acc * 0.5f + k), 0.5f);
      }
      if ((MODE & 4) && warp < 8) {
        tmem_ld16(tmem + (uint32_t((warp & 3) * 32) << 16) + 448 + (warp >> 2) * 16, v);
        tmem_ld_wait();
        acc += v[3];
      }
      if (MODE & 8) {
        const uint32_t a = smem_u32(smem + 150000) + threadIdx.x * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) { sts_f32(a + k * 2048, acc); acc += lds_f32(a + ((k + 3) & 7) * 2048); }
      }
      if (!(MODE & 14)) __nanosleep(100);
    }
    if (acc == 1234.5f) sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem);
}

template <int MODE>
void run(const uint8_t* gsrc, float* sink) {
  long long* d; cudaMalloc(&d, 16);
  auto k = bench<MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int groups = 400;
  k<<<148, 640, 200 * 1024>>>(d, gsrc, sink, groups);
  k<<<148, 640, 200 * 1024>>>(d, gsrc, sink, groups);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("mode %2d (1 bulk loads into smem, 2 MUFU/FMA warps, 4 tcgen05.ld warps, 8 smem ld/st warps): %.1f clk/MMA (%s)\n", MODE,
         (double)h[0] / (24.0 * groups), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  uint8_t* g; cudaMalloc(&g, (size_t)4096 * 8192); cudaMemset(g, 0, (size_t)4096 * 8192);
  float* sink; cudaMalloc(&sink, 4096);
  run<0>(g, sink); run<1>(g, sink); run<2>(g, sink); run<4>(g, sink); run<8>(g, sink); run<15>(g, sink);
  return 0;
}
