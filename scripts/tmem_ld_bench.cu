// Microbenchmark: what does the TMEM read port of an sm_100a SM count — bytes delivered to the register file or
// TMEM columns touched?  (the fcomb kernels are bound by tcgen05.ld of their fp32 accumulators at
// ~62 B/clk/SM; the f16 variants read the same columns with .pack::16b, i.e. half the register bytes.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I probabilistic-multiplanar-unet_b200/csrc \
//        -o /tmp/tmem_ld_bench scripts/tmem_ld_bench.cu && /tmp/tmem_ld_bench
//
// One CTA per SM, NW warps (warp w reads lane quarter w % 4, column block (w / 4) * 32 ...), each warp issues ITER
// loads of 32 TMEM columns back to back (DEPTH loads in flight per tcgen05.wait::ld) and times itself with clock64.
// Reported per variant: cycles per load per warp, register bytes / clk / SM and TMEM-column bytes / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

using namespace pmu::ptx;

template <int VARIANT>
__device__ __forceinline__ void ld32cols(uint32_t taddr, uint32_t (&r)[32]) {
  if (VARIANT == 0) {                       // 32 columns -> 32 registers
    tmem_ld_32x32(taddr, r);
  } else if (VARIANT == 1) {                // 32 columns -> 16 registers (.pack::16b)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
  } else {                                  // 16 columns -> 16 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
  }
}

// VARIANT 0: x32 fp32 (32 cols, 128 B / thread)   1: x16.pack::16b (32 cols, 64 B / thread)   2: x16 (16 cols, 64 B / thread)
template <int VARIANT, int DEPTH>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tptr));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) ld32cols<VARIANT>(base + (uint32_t)((((warp >> 2) * DEPTH + d) * 32) & 511), r[d]);
    tmem_ld_wait();                          // registers are only defined after the wait
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int k = 0; k < (VARIANT == 0 ? 32 : 16); ++k) acc ^= r[d][k];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0xdeadbeefu) sink[0] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tptr);
}

template <int VARIANT, int DEPTH>
static void run(const char* name, int nwarps, int sms, long long* d_cycles, uint32_t* d_sink) {
  const int iters = 4096;
  cudaMemset(d_cycles, 0, sizeof(long long) * sms * 16);
  tmem_ld_kernel<VARIANT, DEPTH><<<sms, nwarps * 32>>>(64, d_cycles, d_sink);          // warm-up
  tmem_ld_kernel<VARIANT, DEPTH><<<sms, nwarps * 32>>>(iters, d_cycles, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s warps %2d: %s\n", name, nwarps, cudaGetErrorString(e)); return; }
  static long long h[148 * 16 * 2];
  cudaMemcpy(h, d_cycles, sizeof(long long) * sms * 16, cudaMemcpyDeviceToHost);
  long long worst = 0;
  for (int i = 0; i < sms * 16; ++i) worst = h[i] > worst ? h[i] : worst;
  const double loads = (double)iters * DEPTH * nwarps;                       // warp-level loads per SM
  const double reg_bytes = loads * 32 * (VARIANT == 0 ? 128 : 64);
  const double col_bytes = loads * 32 * (VARIANT == 2 ? 16 : 32) * 4;
  printf("%-28s warps %2d depth %d: %7.1f clk / load / warp   %6.1f reg B/clk/SM   %6.1f TMEM-column B/clk/SM\n", name, nwarps,
         DEPTH, (double)worst / (iters * DEPTH), reg_bytes / worst, col_bytes / worst);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  if (sms <= 0 || sms > 296) sms = 148;
  long long* d_cycles; uint32_t* d_sink;
  cudaMalloc(&d_cycles, sizeof(long long) * 296 * 16);
  cudaMalloc(&d_sink, 64);
  for (int nw : {4, 8, 16}) {
    run<0, 1>("x32 fp32 (32 col)", nw, sms, d_cycles, d_sink);
    run<0, 2>("x32 fp32 (32 col)", nw, sms, d_cycles, d_sink);
    run<1, 1>("x16.pack::16b (32 col)", nw, sms, d_cycles, d_sink);
    run<1, 2>("x16.pack::16b (32 col)", nw, sms, d_cycles, d_sink);
    run<2, 1>("x16 fp32 (16 col)", nw, sms, d_cycles, d_sink);
    run<2, 2>("x16 fp32 (16 col)", nw, sms, d_cycles, d_sink);
  }
  return 0;
}
