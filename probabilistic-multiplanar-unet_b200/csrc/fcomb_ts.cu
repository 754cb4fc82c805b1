// K3 + K4 fused, TS-mode variant — the fcomb MLP with its activations RESIDENT IN TENSOR MEMORY.
//
// fcomb_tc6.cu is bound by the shared-memory pipe (activation stores + UMMA A-operand fetches, see
// its header).  Here the hidden activations never touch shared memory: the epilogue reads the fp32
// accumulator X from TMEM (tcgen05.ld), applies ReLU + bf16 pack and writes the packed row back to
// TMEM (tcgen05.st) as the A operand Y of the next layer's UMMA ("TS" form: A from TMEM, B = weights
// from smem).  Only the weight tiles (2 KB per K = 16 step) are fetched from shared memory.
//   * one 128-pixel tile per CTA at a time, 16 epilogue warps: warp = (TMEM lane quarter, 16-column
//     quarter), i.e. 4 threads per pixel row, 16 accumulator columns each;
//   * four samples (slots) in flight; TMEM per slot: X = 64 fp32 columns, Y = 40 columns = 64 bf16
//     activations (32 columns) + a K extension of 16 whose first two entries are constant ones, so the
//     bias (bf16 hi + lo in the matching weight columns) rides in the GEMM as a fifth K = 16 step;
//   * layer 0 as in fcomb_tc6.cu: G = W0f f once per tile (SS UMMA from the TMA-loaded feature tile),
//     kept in registers (16 per thread); per sample h0 = relu(G + zb_n) goes straight to Y;
//   * softmax of slot s is done by column quarter s (one thread per pixel), partial sums combined
//     through shared memory at the end of the tile.
// Replaces Fcomb.forward / softmax / the sample loop (probabilistic_unet.py:155-181, eval.py:146-157).
//
// F16 = true (PMU_FCOMB_TS=2, EXPERIMENT, not the default): the per-sample hidden layers run f16 x f16 -> f16.  A dense
// UMMA keeps an f16 accumulator in the low half of a 32-bit TMEM column; `tcgen05.ld ... .pack::16b` returns two
// adjacent columns per register, which is already the packed A-operand layout of the next layer: the epilogue of a
// hidden layer is ld (8 registers) -> max.f16x2 with 0 -> st.  If the TMEM read port is paced by the bytes delivered
// to the register file (64 B/clk, see fcomb_tc6.cu) this halves the dominant cost; if it is paced by the columns
// touched it changes nothing.  Layer 0 (G = W0f f, bf16 features, fp32 accumulate, exact fp32 bias) and the logits
// (fp32 accumulate) are unchanged; the hidden activations carry 11 significand bits instead of 8.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "pmu_common.cuh"
#include "sm100_ptx.cuh"

namespace pmu {

using namespace ptx;

constexpr int FT_F = 64;
constexpr int FT_SLOTS = 4;
constexpr int FT_EPI = 512;                       // 16 epilogue warps
constexpr int FT_THREADS = FT_EPI + 32;           // + issuer warp
constexpr int FT_NS = 16;
constexpr int FT_SLOT_COLS = 104;                 // X 64 + Y 40
constexpr int FT_TILE = 128 * 128;
constexpr int FT_WT = 64 * 128;
constexpr int FT_OFF_W0 = 0;
constexpr int FT_OFF_WM = FT_OFF_W0 + FT_WT;       // 2 mid layers
constexpr int FT_OFF_WL = FT_OFF_WM + 2 * FT_WT;   // head [16][64]
constexpr int FT_OFF_BMT = FT_OFF_WL + 2048;       // bias tiles (k0 = hi, k1 = lo)
constexpr int FT_OFF_BLT = FT_OFF_BMT + 2 * FT_WT;
constexpr int FT_OFF_F = FT_OFF_BLT + 2048;        // feature tile, double buffered
constexpr int FT_OFF_ZB = FT_OFF_F + 2 * FT_TILE;  // fp32 zb[FT_NS][64]
constexpr int FT_OFF_SCR = FT_OFF_ZB + FT_NS * FT_F * 4;       // softmax partials [4 quarters][16][128]
constexpr int FT_OFF_BAR = FT_OFF_SCR + 4 * 16 * 128 * 4;
constexpr int FT_NBAR = 2 * FT_SLOTS + 4;          // ready[slot], acc[slot], tma[2], g, free
constexpr int FT_OFF_TPTR = FT_OFF_BAR + FT_NBAR * 8;
constexpr int FT_SMEM = FT_OFF_TPTR + 16;
static_assert(FT_OFF_F % 1024 == 0 && FT_OFF_BMT % 1024 == 0 && FT_OFF_BLT % 1024 == 0, "operand tiles must be 1024 B aligned");
static_assert(FT_SLOTS * FT_SLOT_COLS <= 512, "TMEM budget");

struct FcombTsParams {
  int N, L, C, nmid, B;
  int64_t HW;
};

__device__ __forceinline__ void ft_st_bf16(uint8_t* tile, int row, int k, float v) {
  *reinterpret_cast<__nv_bfloat16*>(tile + (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2)) = __float2bfloat16(v);
}
template <bool F16>
__device__ __forceinline__ void ft_st_w(uint8_t* tile, int row, int k, float v) {   // weight / bias element of a per-sample layer
  if constexpr (F16)
    *reinterpret_cast<__half*>(tile + (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2)) = __float2half_rn(v);
  else
    ft_st_bf16(tile, row, k, v);
}
template <bool F16>
__device__ __forceinline__ float ft_round_w(float v) {
  if constexpr (F16) return __half2float(__float2half_rn(v));
  else return __bfloat162float(__float2bfloat16(v));
}
// relu(a + b) of two fp32 pairs -> packed f16x2
__device__ __forceinline__ uint32_t ft_add_pack_relu_h(float a0, float a1, float b0, float b1) {
  uint32_t d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t.reg .f32 lo, hi;\n\t"
      "mov.b64 ra, {%1, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {lo, hi}, rd;\n\t"
      "cvt.rn.relu.f16x2.f32 %0, hi, lo;\n\t}"
      : "=r"(d) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
  return d;
}
__device__ __forceinline__ uint32_t ft_relu_h2(uint32_t v) {
  uint32_t d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0u));
  return d;
}
// 32 lanes x 16 columns holding one f16 each (low half) -> 8 registers of f16x2 (column 2j low, 2j + 1 high)
__device__ __forceinline__ void ft_tmem_ld16_pack(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
// kind::f16 instruction descriptor, f16 x f16 operands (formats 0), accumulator f16 (D format 0) or fp32 (1)
__host__ __device__ constexpr uint32_t ft_idesc_f16(int M, int N, bool acc_f32) {
  return ((acc_f32 ? 1u : 0u) << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t ft_pack_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t ft_add_pack_relu(float a0, float a1, float b0, float b1) {
  uint32_t d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t.reg .f32 lo, hi;\n\t"
      "mov.b64 ra, {%1, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {lo, hi}, rd;\n\t"
      "cvt.rn.relu.bf16x2.f32 %0, hi, lo;\n\t}"
      : "=r"(d) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
  return d;
}
__device__ __forceinline__ float4 ft_lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ft_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ft_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ft_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ft_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]   (TS form: the A operand is read from tensor memory)
__device__ __forceinline__ void ft_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one layer on tensor memory: X = Y[128 x 80] * [W | bias]^T   (4 + 1 TS UMMAs; Y columns 32..39 hold the ones)
__device__ __forceinline__ void ft_issue_layer(uint32_t tX, uint32_t tY, uint32_t w_tile, uint32_t b_tile, uint32_t idesc) {
  const uint64_t wd = umma_smem_desc_sw128(w_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) ft_umma_ts(tX, tY + 8 * k, wd + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
  ft_umma_ts(tX, tY + 32, umma_smem_desc_sw128(b_tile), idesc, 1u);
}

template <int CMAX, bool F16>
__global__ void __launch_bounds__(FT_THREADS, 1)
fcomb_ts_kernel(const __grid_constant__ CUtensorMap tmF, const FcombTsParams p, const float* __restrict__ mu,
                const float* __restrict__ sigma, const float* __restrict__ eps, const float* __restrict__ w0,
                const float* __restrict__ b0, const float* __restrict__ wmid, const float* __restrict__ bmid,
                const float* __restrict__ wlast, const float* __restrict__ blast, float* __restrict__ slice_sums) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  uint8_t* sgen = smem_raw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, L = p.L, C = p.C, nmid = p.nmid;
  const int64_t HW = p.HW;
  if ((sbase & 1023u) != 0) __trap();

  auto bar_ready = [&](int s) { return sbase + FT_OFF_BAR + s * 8; };
  auto bar_acc = [&](int s) { return sbase + FT_OFF_BAR + (FT_SLOTS + s) * 8; };
  auto bar_tma = [&](int i) { return sbase + FT_OFF_BAR + (2 * FT_SLOTS + i) * 8; };
  const uint32_t bar_g = sbase + FT_OFF_BAR + (2 * FT_SLOTS + 2) * 8;
  const uint32_t bar_free = sbase + FT_OFF_BAR + (2 * FT_SLOTS + 3) * 8;
  volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(sgen + FT_OFF_TPTR);

  if (tid == 0) {
    prefetch_tensormap(&tmF);
    for (int s = 0; s < FT_SLOTS; ++s) { mbar_init(bar_ready(s), FT_EPI / 32); mbar_init(bar_acc(s), 1); }   // one arrival per warp
    mbar_init(bar_tma(0), 1); mbar_init(bar_tma(1), 1);
    mbar_init(bar_g, 1);
    mbar_init(bar_free, FT_EPI / 32);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc<512>(sbase + FT_OFF_TPTR);
  for (int i = tid; i < FT_OFF_F / 16; i += FT_THREADS) reinterpret_cast<uint4*>(sgen)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < FT_F * FT_F; i += FT_THREADS) {
    const int o = i >> 6, k = i & 63;
    ft_st_bf16(sgen + FT_OFF_W0, o, k, __ldg(w0 + (int64_t)o * (FT_F + L) + k));
    for (int m = 0; m < nmid; ++m) ft_st_w<F16>(sgen + FT_OFF_WM + m * FT_WT, o, k, __ldg(wmid + (int64_t)m * FT_F * FT_F + i));
  }
  for (int i = tid; i < C * FT_F; i += FT_THREADS) ft_st_w<F16>(sgen + FT_OFF_WL, i >> 6, i & 63, __ldg(wlast + i));
  for (int i = tid; i < nmid * FT_F; i += FT_THREADS) {
    const float bv = __ldg(bmid + i);
    const float bh = ft_round_w<F16>(bv);
    ft_st_w<F16>(sgen + FT_OFF_BMT + (i >> 6) * FT_WT, i & 63, 0, bh);
    ft_st_w<F16>(sgen + FT_OFF_BMT + (i >> 6) * FT_WT, i & 63, 1, bv - bh);
  }
  for (int i = tid; i < C; i += FT_THREADS) {
    const float bv = __ldg(blast + i);
    const float bh = ft_round_w<F16>(bv);
    ft_st_w<F16>(sgen + FT_OFF_BLT, i, 0, bh);
    ft_st_w<F16>(sgen + FT_OFF_BLT, i, 1, bv - bh);
  }
  float* zb_s = reinterpret_cast<float*>(sgen + FT_OFF_ZB);
  float* scr = reinterpret_cast<float*>(sgen + FT_OFF_SCR);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tptr;

  // epilogue thread coordinates
  const int q4 = warp & 3, cq = (warp >> 2) & 3;
  const int row = q4 * 32 + lane;
  const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
  if (warp < 16 && cq == 0) {
    // the constant K extension of every slot's A operand: k = 64, 65 -> 1.0 (bf16 pair), k = 66..79 -> 0
    const uint32_t ones[8] = {F16 ? 0x3C003C00u : 0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // (1.0, 1.0) as f16 / bf16 pairs
#pragma unroll
    for (int s = 0; s < FT_SLOTS; ++s) ft_tmem_st8(tmem_base + lane_off + s * FT_SLOT_COLS + 64 + 32, ones);
    ft_tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  const int tiles = (int)((HW + 127) / 128);
  const int64_t total = (int64_t)p.B * tiles;
  const int cta_lo = (int)(total * blockIdx.x / gridDim.x), cta_hi = (int)(total * (blockIdx.x + 1) / gridDim.x);

  uint32_t phr = 0, pha = 0, phg = 0, phf = 0, pht = 0;     // pht: bit i = parity of tma barrier i
  bool first_tile = true;
  uint32_t fbuf = 0;                                        // which F buffer the next tile uses (issuer + everyone in step)

  for (int seg0 = cta_lo; seg0 < cta_hi;) {
    const int b = seg0 / tiles;
    const int seg1 = ((b + 1) * tiles < cta_hi) ? (b + 1) * tiles : cta_hi;
    const int t0 = seg0 - b * tiles, t1 = seg1 - b * tiles;
    for (int n0 = 0; n0 < N; n0 += FT_NS) {
      const int ng = (N - n0 < FT_NS) ? N - n0 : FT_NS;
      __syncthreads();
      for (int i = tid; i < ng * FT_F; i += FT_THREADS) {
        const int n = i >> 6, o = i & 63;
        float s = __ldg(b0 + o);
        for (int l = 0; l < L; ++l) {
          const float z = __fadd_rn(__ldg(mu + (int64_t)b * L + l),
                                    __fmul_rn(__ldg(sigma + (int64_t)b * L + l), __ldg(eps + ((int64_t)b * N + n0 + n) * L + l)));
          s = fmaf(__ldg(w0 + (int64_t)o * (FT_F + L) + FT_F + l), z, s);
        }
        zb_s[i] = s;
      }
      __syncthreads();
      const int rounds = (ng + FT_SLOTS - 1) / FT_SLOTS;

      if (warp == 16) {
        // ============ issuer ============
        // elect.sync (not `lane == 0`): with a provably single active thread the compiler keeps the UMMA descriptors in
        // uniform registers; `lane == 0` costs a 12-instruction waterfall (ELECT / R2UR.BROADCAST / BRA.U.ANY) per UMMA.
        // elect.sync over the full warp always picks the same lane, so the per-thread barrier phases persist.
        if (elect_one()) {
          constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64);                                   // layer 0: bf16 features
          constexpr uint32_t idesc_mid = F16 ? ft_idesc_f16(128, 64, false) : umma_idesc_bf16(128, 64);
          constexpr uint32_t idesc16 = F16 ? ft_idesc_f16(128, 16, true) : umma_idesc_bf16(128, 16);
          const uint32_t sW0 = sbase + FT_OFF_W0, sWM = sbase + FT_OFF_WM, sWL = sbase + FT_OFF_WL;
          const uint32_t sBM = sbase + FT_OFF_BMT, sBL = sbase + FT_OFF_BLT;
          bool f_in_flight = false;
          for (int t = t0; t < t1; ++t) {
            const uint32_t sF = sbase + FT_OFF_F + fbuf * FT_TILE;
            if (!f_in_flight) {
              mbar_arrive_expect_tx(bar_tma(fbuf), FT_TILE);
              tma_load_2d(sF, &tmF, bar_tma(fbuf), 0, (int)((int64_t)b * HW + (int64_t)t * 128));
            }
            f_in_flight = false;
            if (t + 1 < t1) {           // prefetch the next tile's features into the other buffer
              mbar_arrive_expect_tx(bar_tma(fbuf ^ 1u), FT_TILE);
              tma_load_2d(sbase + FT_OFF_F + (fbuf ^ 1u) * FT_TILE, &tmF, bar_tma(fbuf ^ 1u), 0,
                          (int)((int64_t)b * HW + (int64_t)(t + 1) * 128));
              f_in_flight = true;
            }
            if (!first_tile) { mbar_wait(bar_free, phf); phf ^= 1u; }     // slot 0's X is free again
            first_tile = false;
            mbar_wait(bar_tma(fbuf), (pht >> fbuf) & 1u); pht ^= 1u << fbuf;
            tcgen05_fence_after();
            {   // G = F W0f^T -> slot 0's X (SS form)
              const uint64_t ad = umma_smem_desc_sw128(sF), wd = umma_smem_desc_sw128(sW0);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 0, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc64, (uint32_t)(k != 0));
            }
            umma_commit(bar_g);
            fbuf ^= 1u;
            for (int r = 0; r < rounds; ++r) {
              for (int layer = 1; layer <= nmid + 1; ++layer) {
#pragma unroll
                for (int s = 0; s < FT_SLOTS; ++s) {
                  if (r * FT_SLOTS + s >= ng) continue;
                  const uint32_t tX = tmem_base + s * FT_SLOT_COLS, tY = tX + 64;
                  mbar_wait(bar_ready(s), (phr >> s) & 1u); phr ^= 1u << s;
                  tcgen05_fence_after();
                  if (layer <= nmid) ft_issue_layer(tX, tY, sWM + (layer - 1) * FT_WT, sBM + (layer - 1) * FT_WT, idesc_mid);
                  else ft_issue_layer(tX, tY, sWL, sBL, idesc16);
                  umma_commit(bar_acc(s));
                }
              }
            }
          }
        }
        __syncwarp();
        // keep fbuf in step for the (unused) other lanes: only lane 0's copy matters
      } else {
        // ============ epilogue warps: (lane quarter q4, column quarter cq) ============
        const uint32_t tbase = tmem_base + lane_off;
        const uint32_t sZB = sbase + FT_OFF_ZB + cq * 16 * 4;
        for (int t = t0; t < t1; ++t) {
          const int64_t pix = (int64_t)t * 128 + row;
          float s1[CMAX], s2[CMAX];
#pragma unroll
          for (int c = 0; c < CMAX; ++c) s1[c] = s2[c] = 0.f;
          uint32_t G[16];
          mbar_wait(bar_g, phg); phg ^= 1u;
          tcgen05_fence_after();
          ft_tmem_ld16(tbase + cq * 16, G);
          tmem_ld_wait();
          for (int r = 0; r < rounds; ++r) {
            // ---- layer 0: h0 = relu(G + zb_n) -> Y (bf16 pairs, 8 columns per thread) ----
#pragma unroll
            for (int s = 0; s < FT_SLOTS; ++s) {
              const int n = r * FT_SLOTS + s;
              if (n >= ng) continue;
              const uint32_t zb = sZB + n * FT_F * 4;
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 z = ft_lds128f(zb + j * 16);
                if constexpr (F16) {
                  pk[2 * j] = ft_add_pack_relu_h(__uint_as_float(G[4 * j]), __uint_as_float(G[4 * j + 1]), z.x, z.y);
                  pk[2 * j + 1] = ft_add_pack_relu_h(__uint_as_float(G[4 * j + 2]), __uint_as_float(G[4 * j + 3]), z.z, z.w);
                } else {
                  pk[2 * j] = ft_add_pack_relu(__uint_as_float(G[4 * j]), __uint_as_float(G[4 * j + 1]), z.x, z.y);
                  pk[2 * j + 1] = ft_add_pack_relu(__uint_as_float(G[4 * j + 2]), __uint_as_float(G[4 * j + 3]), z.z, z.w);
                }
              }
              ft_tmem_st8(tbase + s * FT_SLOT_COLS + 64 + cq * 8, pk);
              ft_tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_ready(s));      // one arrival per warp (barrier count = 16)
            }
            // ---- mid layers: X -> relu -> bf16 -> Y ----
            for (int layer = 1; layer <= nmid; ++layer) {
#pragma unroll
              for (int s = 0; s < FT_SLOTS; ++s) {
                if (r * FT_SLOTS + s >= ng) continue;
                mbar_wait(bar_acc(s), (pha >> s) & 1u); pha ^= 1u << s;
                tcgen05_fence_after();
                uint32_t pk[8];
                if constexpr (F16) {
                  ft_tmem_ld16_pack(tbase + s * FT_SLOT_COLS + cq * 16, pk);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 8; ++j) pk[j] = ft_relu_h2(pk[j]);
                } else {
                  uint32_t rr[16];
                  ft_tmem_ld16(tbase + s * FT_SLOT_COLS + cq * 16, rr);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 8; ++j) pk[j] = ft_pack_relu(__uint_as_float(rr[2 * j]), __uint_as_float(rr[2 * j + 1]));
                }
                ft_tmem_st8(tbase + s * FT_SLOT_COLS + 64 + cq * 8, pk);
                ft_tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready(s));
              }
            }
            // ---- head: softmax of slot s by column quarter s ----
#pragma unroll
            for (int s = 0; s < FT_SLOTS; ++s) {
              if (r * FT_SLOTS + s >= ng) continue;
              mbar_wait(bar_acc(s), (pha >> s) & 1u); pha ^= 1u << s;
              tcgen05_fence_after();
              if (cq == s) {
                uint32_t hr[8];
                ft_tmem_ld8(tbase + s * FT_SLOT_COLS, hr);
                tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) if (c < C) mx = fmaxf(mx, __uint_as_float(hr[c]));
                float e[CMAX], den = 0.f;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) { e[c] = (c < C) ? __expf(__uint_as_float(hr[c]) - mx) : 0.f; den += e[c]; }
                const float inv = __fdividef(1.f, den);
#pragma unroll
                for (int c = 0; c < CMAX; ++c) { const float pr_ = e[c] * inv; s1[c] += pr_; s2[c] = fmaf(pr_, pr_, s2[c]); }
              }
            }
          }
          // ---- tile done ----
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_free);
#pragma unroll
          for (int c = 0; c < CMAX; ++c) { scr[(cq * 16 + 2 * c) * 128 + row] = s1[c]; scr[(cq * 16 + 2 * c + 1) * 128 + row] = s2[c]; }
          named_bar_sync(1, FT_EPI);
          if (cq == 0 && pix < HW) {
            float* o1 = slice_sums + ((int64_t)b * 2 + 0) * C * HW + pix;
            float* o2 = slice_sums + ((int64_t)b * 2 + 1) * C * HW + pix;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
              if (c < C) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) { a1 += scr[(qq * 16 + 2 * c) * 128 + row]; a2 += scr[(qq * 16 + 2 * c + 1) * 128 + row]; }
                if (n0 == 0) { o1[(int64_t)c * HW] = a1; o2[(int64_t)c * HW] = a2; }
                else { o1[(int64_t)c * HW] += a1; o2[(int64_t)c * HW] += a2; }
              }
          }
          named_bar_sync(1, FT_EPI);
        }
      }
    }
    seg0 = seg1;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem_base);
}

static PFN_cuTensorMapEncodeTiled_v12000 ft_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace pmu

using namespace pmu;

// called by pmu_fcomb_softmax_accum_bf16 (fcomb_tc6.cu) after argument checks
extern "C" int pmu_fcomb_softmax_accum_bf16_ts(const void* feat, const float* mu, const float* sigma, const float* eps,
                                               const float* w0, const float* b0, const float* wmid, const float* bmid,
                                               const float* wlast, const float* blast, float* slice_sums, int B, int N,
                                               int L, int C, int nl, int64_t HW, void* stream) {
  auto fn = ft_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PMU_ERR_CUDA; }
  CUtensorMap tmF;
  cuuint64_t dims[2] = {(cuuint64_t)FT_F, (cuuint64_t)((int64_t)B * HW)};
  cuuint64_t strides[1] = {(cuuint64_t)FT_F * 2};
  cuuint32_t box[2] = {(cuuint32_t)FT_F, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tmF, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(feat), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(features) failed: %d", (int)r); return PMU_ERR_CUDA; }
  FcombTsParams p;
  p.N = N; p.L = L; p.C = C; p.nmid = nl - 2; p.HW = HW; p.B = B;
  const int64_t tiles = (HW + 127) / 128;
  const int64_t total = (int64_t)B * tiles;
  const unsigned grid = (unsigned)std::min<int64_t>(total, sm_count());
  // PMU_FCOMB_TS=2 (the dispatcher in fcomb_tc6.cu only comes here for a non-zero value): f16 hidden layers (experiment)
  const char* ts_env = getenv("PMU_FCOMB_TS");
  const bool f16 = ts_env && atoi(ts_env) == 2;
  auto launch = [&](auto kern) -> int {
    PMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    kern<<<grid, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(tmF, p, mu, sigma, eps, w0, b0, wmid, bmid, wlast, blast, slice_sums);
    PMU_LAUNCH_CHECK();
    return PMU_OK;
  };
  if (C <= 4) return f16 ? launch(fcomb_ts_kernel<4, true>) : launch(fcomb_ts_kernel<4, false>);
  return f16 ? launch(fcomb_ts_kernel<8, true>) : launch(fcomb_ts_kernel<8, false>);
}
