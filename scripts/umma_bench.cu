// Microbenchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16 -> fp32) as a function of N, for the SS form
// (A and B from shared memory) and the TS form (A from tensor memory), issued back to back by one thread.
// Answers whether an N = 64 UMMA costs 32 cycles (the M*N/256 floor) or more, i.e. what bounds the fcomb kernels
// and the 64-cout convolutions.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I probabilistic-multiplanar-unet_b200/csrc \
//        -o /tmp/umma_bench scripts/umma_bench.cu && /tmp/umma_bench
//
// One CTA per SM (148 CTAs, so the figure is taken under chip-wide load); 64 threads: warp 0 allocates TMEM, one
// elected thread issues REPS groups of `chain` UMMAs (a group accumulates into one of `ndst` accumulators, the
// groups rotate over them), commits once at the end and waits.  Operand tiles are whatever shared memory holds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

using namespace pmu::ptx;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// form 0 = SS, 1 = TS; a_stride / b_stride: byte distance between the operand tiles of successive UMMAs of a chain
// (0 = the same 16-wide K slice again, 32 = walk the four K slices of a 128 B swizzle row like a real K loop)
template <int form, int chain>
__global__ void __launch_bounds__(64, 1) umma_kernel(int N, int reps, int ndst, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += 64) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tptr));
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tptr;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t adesc = umma_smem_desc_sw128(sbase), bdesc = umma_smem_desc_sw128(sbase + 16384);
    const uint32_t ta = tmem + 448;                       // A operand columns for the TS form (64 columns: 4 K slices + spare)
    const int dst_cols = (N < 32) ? 32 : N;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (uint32_t)((r % ndst) * dst_cols);
#pragma unroll
      for (int k = 0; k < chain; ++k) {
        if (form == 0) umma_bf16(d, adesc + (uint64_t)(2 * (k & 3)), bdesc + (uint64_t)(2 * (k & 3)), idesc, (uint32_t)(k != 0));
        else umma_ts(d, ta + 8 * (k & 3), bdesc + (uint64_t)(2 * (k & 3)), idesc, (uint32_t)(k != 0));
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, 8);
  const int Ns[] = {8, 16, 32, 64, 128, 256};
  printf("M = 128, K = 16 per UMMA, bf16 x bf16 -> fp32, one issuing thread per SM, 148 CTAs\n");
  for (int cfg = 0; cfg < 4; ++cfg)
      for (int ni = 0; ni < 6; ++ni) {
        const int form = cfg >> 1, chain = (cfg & 1) ? 36 : 5;
        const int N = Ns[ni];
        const int ndst = (N >= 256) ? 1 : (N >= 128 ? 2 : 4);
        const int reps = 4096 / chain * 4;
        auto run = [&](auto kern) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
          kern<<<148, 64, 64 * 1024>>>(N, 8, ndst, d_cycles);     // warm-up
          kern<<<148, 64, 64 * 1024>>>(N, reps, ndst, d_cycles);
        };
        if (cfg == 0) run(umma_kernel<0, 5>); else if (cfg == 1) run(umma_kernel<0, 36>);
        else if (cfg == 2) run(umma_kernel<1, 5>); else run(umma_kernel<1, 36>);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d: %s\n", form ? "TS" : "SS", N, cudaGetErrorString(e)); return 1; }
        long long c = 0;
        cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / ((double)reps * chain);
        printf("%s  N = %3d  chain %2d (x%d accumulators): %7.1f clk / UMMA   %7.0f MAC/clk/SM   (floor M*N/256 = %d clk)\n",
               form ? "TS" : "SS", N, chain, ndst, per, 128.0 * N * 16 / per, 128 * N / 256);
      }
  return 0;
}
