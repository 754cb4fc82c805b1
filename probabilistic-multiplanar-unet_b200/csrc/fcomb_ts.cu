// K3 + K4 fused — the fcomb MLP over N latent samples with its hidden activations RESIDENT IN TENSOR
// MEMORY (TS-form UMMAs: A operand from TMEM, B = weights from shared memory), softmax, per-pixel
// sum / sum-of-squares.  Per-sample logits never reach HBM.
//
// Replaces Fcomb.forward (probabilistic_unet.py:155-181) called once per sample from
// ProbabilisticUnet.sample (:225-240), the softmax of eval.py:157 and the sample loop of
// eval.py:146-154.
//
// What the round-2 measurements said about the two earlier tcgen05 versions (both ~1130 cycles per 128-pixel
// tile-sample, tensor pipe ~32 % busy):
//   * the TMEM read port is NOT the limit: scripts/tmem_ld_bench.cu measures 205 / 365-413 / 470-480 B/clk/SM with
//     4 / 8 / 16 warps (register bytes; .pack::16b loads deliver the same register bytes), against the ~62 B/clk the
//     kernels used (profiles/r02_experiments.txt);
//   * the SS form (activations through shared memory) is bound by the shared-memory pipe: per tile-sample 10 N = 64
//     UMMAs x 6 KB of operand reads + 5 head UMMAs x 4.25 KB + 48 KB of activation stores = 129 KB = 1008 cycles of
//     the 128 B/clk pipe;
//   * the first TS form kept all 16 epilogue warps in lock-step on ONE slot at a time (16 columns per thread), so the
//     per-slot chain  barrier wake-up -> tcgen05.ld -> wait -> pack -> tcgen05.st -> wait -> arrive  (~300 cycles)
//     was paid 4 slots x 3 layers = 12 times per round of four samples, one after the other.
// This version: the 16 epilogue warps form two groups of 8 (TMEM lane quarter x column half: 32 accumulator columns
// per thread); a group owns two of the four sample slots and walks them alternately, so while one slot's epilogue
// runs the other slot's UMMAs execute, and the two groups run independently of each other.  Per slot and layer the
// critical path is  commit -> wake-up -> ld.x32 -> 16 cvt -> st.x16 -> arrive -> issue, four such chains in flight.
//   * layer 0 leaves the per-sample chain: h0_n = relu(W0f f + zb_n), zb_n = W0z z_n + b0.  G = W0f f is ONE SS-form
//     UMMA group per tile into its own TMEM region (so the next tile's G is computed while the current tile's
//     samples run), read once into registers; per sample layer 0 is 16 add.f32x2 + 16 cvt.rn.relu.bf16x2 per thread;
//   * the constant biases ride in the GEMM: every layer's K is extended by one K = 16 step whose A columns are
//     constant ones (8 TMEM columns per slot) and whose B tile holds the bias split into bf16 hi + lo;
//   * TMEM: 4 slots x (X 64 fp32 columns + Y 40 columns = 64 bf16 activations + the ones extension) + G 64 = 480;
//   * the head's softmax of a slot is done by the column half that matches the slot's parity (one thread per pixel),
//     the four partial (sum, sum^2) sets of a pixel are combined through shared memory at the end of the tile.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "pmu_common.cuh"
#include "h16.cuh"
#include "sm100_ptx.cuh"

namespace pmu {

using namespace ptx;

constexpr int F2_F = 64;
constexpr int F2_SLOTS = 4;
constexpr int F2_EPI = 512;                       // 16 epilogue warps
constexpr int F2_THREADS = F2_EPI + 4 * 32;       // + one issuer warp per sample slot
constexpr int F2_NS = 16;                         // samples per group of zb vectors
constexpr int F2_SLOT_COLS = 104;                 // X 64 + Y 40
constexpr int F2_G_COL = F2_SLOTS * F2_SLOT_COLS; // 416: G = W0f f of the current / next tile
constexpr int F2_MAXL = 16;
constexpr int F2_MAXC = 8;
constexpr int F2_TILE = 128 * 128;
constexpr int F2_WT = 64 * 128;
constexpr int F2_OFF_W0 = 0;
constexpr int F2_MAXMID = 4;                       // hidden layers after layer 0 (no_convs_fcomb <= 6)
constexpr int F2_OFF_WM = F2_OFF_W0 + F2_WT;       // mid layers
constexpr int F2_OFF_WL = F2_OFF_WM + F2_MAXMID * F2_WT;   // head [16][64]
constexpr int F2_OFF_BMT = F2_OFF_WL + 2048;       // bias tiles (k0 = hi, k1 = lo)
constexpr int F2_OFF_BLT = F2_OFF_BMT + F2_MAXMID * F2_WT;
constexpr int F2_OFF_F = F2_OFF_BLT + 2048;        // feature tile, double buffered
constexpr int F2_OFF_ZB = F2_OFF_F + 2 * F2_TILE;  // fp32 zb[F2_NS][64]
constexpr int F2_SCR_BYTES = 4 * 16 * 128 * 4;     // softmax partials [4 (group, half)][16][128]
constexpr int F2_OFF_SCR = F2_OFF_ZB + F2_NS * F2_F * 4;       // x 2 (alternating tiles)
constexpr int F2_OFF_BAR = F2_OFF_SCR + 2 * F2_SCR_BYTES;
constexpr int F2_NBAR = 2 * F2_SLOTS + 4;          // ready[slot], acc[slot], tma[2], g_full, g_free
constexpr int F2_OFF_TPTR = F2_OFF_BAR + F2_NBAR * 8;
constexpr int F2_SMEM = F2_OFF_TPTR + 16;
static_assert(F2_OFF_F % 1024 == 0 && F2_OFF_BMT % 1024 == 0 && F2_OFF_BLT % 1024 == 0, "operand tiles must be 1024 B aligned");
static_assert(F2_G_COL + 64 <= 512, "TMEM budget");
static_assert(F2_SMEM <= 227 * 1024, "shared memory budget");

#ifdef F2_TRACE
// scripts/fcomb_trace.cu: clock stamps of CTA 0's issuer and of one epilogue warp per slot group, kept in shared memory
// (one CS2R + one STS per stamp) and copied out at the end: word = id << 24 | clock[23:0]
__device__ uint32_t* f2_trace_buf = nullptr;        // [3 recorders][1024]
constexpr int F2_TRACE_OFF = F2_SMEM;
constexpr int F2_SMEM_TOTAL = F2_SMEM + 3 * 1024 * 4;
#define F2_T(id) do { if (trace_rec >= 0 && trace_idx < 1024) { trace_s[trace_rec * 1024 + trace_idx++] = ((uint32_t)(id) << 24) | ((uint32_t)clock() & 0xFFFFFFu); } } while (0)
#else
constexpr int F2_SMEM_TOTAL = F2_SMEM;
#define F2_T(id) do { } while (0)
#endif

struct FcombTsParams {
  int N, L, C, nmid, B;
  int64_t HW;
};

// weight / bias tiles in the activations' 16-bit format (h16.cuh)
template <bool F16>
__device__ __forceinline__ void f2_st_w(uint8_t* tile, int row, int k, float v) {
  *reinterpret_cast<uint16_t*>(tile + (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2)) = cvt16<F16>(v);
}
__device__ __forceinline__ float4 f2_lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void f2_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void f2_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void f2_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void f2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]   (TS form: the A operand is read from tensor memory)
__device__ __forceinline__ void f2_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one layer on tensor memory: X = Y[128 x 80] * [W | bias]^T   (4 + 1 TS UMMAs; Y columns 32..39 hold the ones)
__device__ __forceinline__ void f2_issue_layer(uint32_t tX, uint32_t tY, uint32_t w_tile, uint32_t b_tile, uint32_t idesc) {
  const uint64_t wd = umma_smem_desc_sw128(w_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) f2_umma_ts(tX, tY + 8 * k, wd + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
  f2_umma_ts(tX, tY + 32, umma_smem_desc_sw128(b_tile), idesc, 1u);
}

// bounded wait without the clock: a pipeline bug traps after ~2^24 wake-ups instead of hanging the GPU
__device__ __forceinline__ void f2_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void f2_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int CMAX, bool F16>
__global__ void __launch_bounds__(F2_THREADS, 1)
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

  auto bar_ready = [&](int s) { return sbase + F2_OFF_BAR + s * 8; };
  auto bar_acc = [&](int s) { return sbase + F2_OFF_BAR + (F2_SLOTS + s) * 8; };
  auto bar_tma = [&](int i) { return sbase + F2_OFF_BAR + (2 * F2_SLOTS + i) * 8; };
  const uint32_t bar_g = sbase + F2_OFF_BAR + (2 * F2_SLOTS + 2) * 8;
  const uint32_t bar_gfree = sbase + F2_OFF_BAR + (2 * F2_SLOTS + 3) * 8;
  volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(sgen + F2_OFF_TPTR);

  if (tid == 0) {
    prefetch_tensormap(&tmF);
    for (int s = 0; s < F2_SLOTS; ++s) { mbar_init(bar_ready(s), 4); mbar_init(bar_acc(s), 1); }   // 4 warps own a slot
    mbar_init(bar_tma(0), 1); mbar_init(bar_tma(1), 1);
    mbar_init(bar_g, 1);
    mbar_init(bar_gfree, F2_EPI / 32);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc<512>(sbase + F2_OFF_TPTR);
  for (int i = tid; i < F2_OFF_F / 16; i += F2_THREADS) reinterpret_cast<uint4*>(sgen)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < F2_F * F2_F; i += F2_THREADS) {
    const int o = i >> 6, k = i & 63;
    f2_st_w<F16>(sgen + F2_OFF_W0, o, k, __ldg(w0 + (int64_t)o * (F2_F + L) + k));
    for (int m = 0; m < nmid; ++m) f2_st_w<F16>(sgen + F2_OFF_WM + m * F2_WT, o, k, __ldg(wmid + (int64_t)m * F2_F * F2_F + i));
  }
  for (int i = tid; i < C * F2_F; i += F2_THREADS) f2_st_w<F16>(sgen + F2_OFF_WL, i >> 6, i & 63, __ldg(wlast + i));
  for (int i = tid; i < nmid * F2_F; i += F2_THREADS) {
    const float bv = __ldg(bmid + i);
    const float bh = cvt16_to_f32<F16>(cvt16<F16>(bv));
    f2_st_w<F16>(sgen + F2_OFF_BMT + (i >> 6) * F2_WT, i & 63, 0, bh);
    f2_st_w<F16>(sgen + F2_OFF_BMT + (i >> 6) * F2_WT, i & 63, 1, bv - bh);
  }
  for (int i = tid; i < C; i += F2_THREADS) {
    const float bv = __ldg(blast + i);
    const float bh = cvt16_to_f32<F16>(cvt16<F16>(bv));
    f2_st_w<F16>(sgen + F2_OFF_BLT, i, 0, bh);
    f2_st_w<F16>(sgen + F2_OFF_BLT, i, 1, bv - bh);
  }
  float* zb_s = reinterpret_cast<float*>(sgen + F2_OFF_ZB);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tptr;

  // epilogue thread coordinates: TMEM lane quarter q4, sample slot sl (4 warps own a slot); thread = one pixel row
  const int q4 = warp & 3, sl = (warp >> 2) & 3;
  const int row = q4 * 32 + lane;
  const uint32_t tX = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(sl * F2_SLOT_COLS);   // accumulator of my slot
  const uint32_t tY = tX + 64;                                                                    // its A operand
  const uint32_t tG = tmem_base + ((uint32_t)(q4 * 32) << 16) + F2_G_COL;
  if (warp < 16) {
    // the constant K extension of the slot's A operand: k = 64, 65 -> 1.0 (bf16 pair), k = 66..79 -> 0
    const uint32_t ones[8] = {F16 ? 0x3C003C00u : 0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // (1.0, 1.0)
    f2_tmem_st8(tY + 32, ones);
    f2_tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  const int tiles = (int)((HW + 127) / 128);
  const int64_t total = (int64_t)p.B * tiles;
  const int cta_lo = (int)(total * blockIdx.x / gridDim.x), cta_hi = (int)(total * (blockIdx.x + 1) / gridDim.x);

  uint32_t phr = 0, pha = 0, phg = 0, phf = 0, pht = 0;     // barrier phase parities (issuer: bit = slot / buffer)
  uint32_t tile_ctr = 0;                                    // scratch buffer selector (epilogue warps)
#ifdef F2_TRACE
  int trace_idx = 0;
  volatile uint32_t* trace_s = reinterpret_cast<volatile uint32_t*>(sgen + F2_TRACE_OFF);
  const int trace_rec = (blockIdx.x != 0) ? -1 : (tid == 0 ? 0 : tid == 256 ? 1 : (warp == 16 && elect_one()) ? 2 : -1);
#endif

  for (int seg0 = cta_lo; seg0 < cta_hi;) {
    const int b = seg0 / tiles;
    const int seg1 = ((b + 1) * tiles < cta_hi) ? (b + 1) * tiles : cta_hi;
    const int t0 = seg0 - b * tiles, t1 = seg1 - b * tiles;
    for (int n0 = 0; n0 < N; n0 += F2_NS) {
      const int ng = (N - n0 < F2_NS) ? N - n0 : F2_NS;
      // ---- per-sample layer-0 bias vectors zb_n = W0z z_n + b0 of this slice / sample group (fp32) ----
      __syncthreads();
      for (int i = tid; i < ng * F2_F; i += F2_THREADS) {
        const int n = i >> 6, o = i & 63;
        float s = __ldg(b0 + o);
        for (int l = 0; l < L; ++l) {
          // z = mu + sigma * eps   (Normal.rsample, probabilistic_unet.py:233)
          const float z = __fadd_rn(__ldg(mu + (int64_t)b * L + l),
                                    __fmul_rn(__ldg(sigma + (int64_t)b * L + l), __ldg(eps + ((int64_t)b * N + n0 + n) * L + l)));
          s = fmaf(__ldg(w0 + (int64_t)o * (F2_F + L) + F2_F + l), z, s);
        }
        zb_s[i] = s;
      }
      __syncthreads();
      const int rounds = (ng + F2_SLOTS - 1) / F2_SLOTS;

      if (warp >= 16) {
        // ============ issuers: warp 16 + s owns sample slot s (one elected thread each: a single issuer spends ~110
        // cycles in every barrier wait with nothing queued behind it, which left the tensor pipe idle half of the time);
        // warp 16 also loads the feature tiles (TMA) and issues the per-tile G = W0f f ============
        if (elect_one()) {
          const int s = warp - 16;
          constexpr uint32_t idesc64 = F16 ? umma_idesc_f16(128, 64) : umma_idesc_bf16(128, 64);
          constexpr uint32_t idesc16 = F16 ? umma_idesc_f16(128, 16) : umma_idesc_bf16(128, 16);
          const uint32_t sW0 = sbase + F2_OFF_W0, sWM = sbase + F2_OFF_WM, sWL = sbase + F2_OFF_WL;
          const uint32_t sBM = sbase + F2_OFF_BMT, sBL = sbase + F2_OFF_BLT;
          const uint32_t sX = tmem_base + s * F2_SLOT_COLS, sY = sX + 64;
          const uint32_t b_ready = bar_ready(s), b_acc = bar_acc(s);
          auto load_f = [&](int t, uint32_t buf) {
            mbar_arrive_expect_tx(bar_tma(buf), F2_TILE);
            tma_load_2d(sbase + F2_OFF_F + buf * F2_TILE, &tmF, bar_tma(buf), 0, (int)((int64_t)b * HW + (int64_t)t * 128));
          };
          auto issue_g = [&](uint32_t buf) {   // G = F W0f^T -> the G columns (SS form)
            f2_wait(bar_tma(buf), (pht >> buf) & 1u); pht ^= 1u << buf;
            tcgen05_fence_after();
            const uint64_t ad = umma_smem_desc_sw128(sbase + F2_OFF_F + buf * F2_TILE), wd = umma_smem_desc_sw128(sW0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + F2_G_COL, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc64, (uint32_t)(k != 0));
            umma_commit(bar_g);
          };
          if (s == 0) {
            load_f(t0, 0);
            if (t0 + 1 < t1) load_f(t0 + 1, 1);
            issue_g(0);                                      // (the previous block's G was released before the __syncthreads)
          }
          for (int t = t0; t < t1; ++t) {
            const uint32_t buf = (uint32_t)(t - t0) & 1u;
            for (int r = 0; r < rounds; ++r) {
              const bool live = r * F2_SLOTS + s < ng;
              for (int layer = 1; layer <= nmid + 1; ++layer) {
                if (live) {
                  F2_T(100 + s);
                  f2_wait(b_ready, phr); phr ^= 1u;
                  tcgen05_fence_after();
                  F2_T(110 + s);
                  if (layer <= nmid) f2_issue_layer(sX, sY, sWM + (layer - 1) * F2_WT, sBM + (layer - 1) * F2_WT, idesc64);
                  else f2_issue_layer(sX, sY, sWL, sBL, idesc16);
                  umma_commit(b_acc);
                  F2_T(120 + s);
                }
                if (s == 0 && r == rounds - 1 && layer == 1) {
                  // once every epilogue warp has read this tile's G for its last sample: compute the next tile's G (it
                  // overlaps the rest of the last round) and refill the feature buffer the previous G used
                  f2_wait(bar_gfree, phf); phf ^= 1u;
                  tcgen05_fence_after();
                  if (t + 1 < t1) {
                    issue_g(buf ^ 1u);
                    if (t + 2 < t1) load_f(t + 2, buf);
                  }
                }
              }
            }
          }
        }
        __syncwarp();
      } else {
        // ============ epilogue warps ============
        const uint32_t sZB = sbase + F2_OFF_ZB;
        const uint32_t b_acc = bar_acc(sl), b_ready = bar_ready(sl);
        const int last_live = (sl < ng) ? (ng - 1 - sl) / F2_SLOTS : -1;    // round of my slot's last sample in this group
        for (int t = t0; t < t1; ++t, ++tile_ctr) {
          const int64_t pix = (int64_t)t * 128 + row;
          float s1[CMAX], s2[CMAX];
#pragma unroll
          for (int c = 0; c < CMAX; ++c) s1[c] = s2[c] = 0.f;
          f2_wait(bar_g, phg); phg ^= 1u;                    // this tile's G = W0f f is in tensor memory
          tcgen05_fence_after();
          if (last_live < 0) { __syncwarp(); if (lane == 0) mbar_arrive(bar_gfree); }

          auto softmax_acc = [&](const uint32_t (&hr)[8]) {
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c < C) mx = fmaxf(mx, __uint_as_float(hr[c]));
            float e[CMAX], den = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) { e[c] = (c < C) ? __expf(__uint_as_float(hr[c]) - mx) : 0.f; den += e[c]; }
            const float inv = __fdividef(1.f, den);
#pragma unroll
            for (int c = 0; c < CMAX; ++c) { const float pr_ = e[c] * inv; s1[c] += pr_; s2[c] = fmaf(pr_, pr_, s2[c]); }
          };

          bool pend = false;                                 // my slot has a head (logits) outstanding
          for (int r = 0; r <= rounds; ++r) {
            const int n = r * F2_SLOTS + sl;
            const bool live = (r < rounds) && (n < ng);
            uint32_t hr[8];
            if (pend) {
              F2_T(70);
              f2_wait(b_acc, pha); pha ^= 1u;                // head UMMAs done: logits in X, Y free
              tcgen05_fence_after();
              f2_tmem_ld8(tX, hr);
              tmem_ld_wait();
            }
            if (live) {
              // layer 0: h0 = relu(G + zb_n) -> Y (bf16 pairs), two halves of 32 columns
              const uint32_t zb = sZB + n * F2_F * 4;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t g[32], pk[16];
                tmem_ld_32x32(tG + h * 32, g);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 z = f2_lds128f(zb + h * 128 + c * 16);
                  pk[2 * c] = add_pack16<true, F16>(__uint_as_float(g[4 * c]), __uint_as_float(g[4 * c + 1]), z.x, z.y);
                  pk[2 * c + 1] = add_pack16<true, F16>(__uint_as_float(g[4 * c + 2]), __uint_as_float(g[4 * c + 3]), z.z, z.w);
                }
                f2_tmem_st16(tY + h * 16, pk);
              }
              f2_tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(b_ready);
                if (r == last_live) mbar_arrive(bar_gfree);  // my last read of this tile's G
              }
              F2_T(80);
            }
            if (pend) softmax_acc(hr);
            pend = live;
            if (!live) break;
            // ---- mid layers: X -> relu -> bf16 -> Y ----
            for (int layer = 1; layer <= nmid; ++layer) {
              F2_T(10);
              f2_wait(b_acc, pha); pha ^= 1u;
              tcgen05_fence_after();
              F2_T(20);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t rr[32], pk[16];
                tmem_ld_32x32(tX + h * 32, rr);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) pk[c] = pack16_relu<F16>(__uint_as_float(rr[2 * c]), __uint_as_float(rr[2 * c + 1]));
                f2_tmem_st16(tY + h * 16, pk);
              }
              F2_T(40);
              f2_tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(b_ready);
              F2_T(60);
            }
          }
          // ---- tile done: combine the four slots' partial sums of a pixel ----
          float* scr = reinterpret_cast<float*>(sgen + F2_OFF_SCR + (tile_ctr & 1u) * F2_SCR_BYTES);
#pragma unroll
          for (int c = 0; c < CMAX; ++c) { scr[(sl * 16 + 2 * c) * 128 + row] = s1[c]; scr[(sl * 16 + 2 * c + 1) * 128 + row] = s2[c]; }
          named_bar_sync(1, F2_EPI);
          if (sl == 0 && pix < HW) {
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
        }
      }
    }
    seg0 = seg1;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem_base);
#ifdef F2_TRACE
  if (trace_rec >= 0 && f2_trace_buf) {
    for (int i = 0; i < 1024; ++i) f2_trace_buf[trace_rec * 1024 + i] = (i < trace_idx) ? trace_s[trace_rec * 1024 + i] : 0u;
  }
#endif
}

static PFN_cuTensorMapEncodeTiled_v12000 f2_encode_fn() {
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

extern "C" int pmu_fcomb_softmax_accum_bf16(const void* feat, const float* mu, const float* sigma,
                                            const float* eps, const float* w0, const float* b0,
                                            const float* wmid, const float* bmid, const float* wlast,
                                            const float* blast, float* slice_sums, int B, int N, int L,
                                            int C, int nl, int64_t HW, int f16, void* stream) {
  PMU_CHECK_ARG(feat && mu && sigma && eps && w0 && b0 && wlast && blast && slice_sums,
                "pmu_fcomb_softmax_accum_bf16: null pointer");
  PMU_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && HW > 0, "pmu_fcomb_softmax_accum_bf16: bad shape");
  PMU_CHECK_ARG(nl >= 2 && (nl == 2 || (wmid && bmid)), "pmu_fcomb_softmax_accum_bf16: no_convs_fcomb >= 2; mid weights needed for > 2");
  PMU_CHECK_SUPPORTED(L >= 1 && L <= F2_MAXL && C >= 1 && C <= F2_MAXC, "pmu_fcomb_softmax_accum_bf16: needs L <= 16, C <= 8 (got L=%d C=%d)", L, C);
  PMU_CHECK_SUPPORTED(nl - 2 <= F2_MAXMID, "pmu_fcomb_softmax_accum_bf16: no_convs_fcomb <= %d on the tensor-core path (got %d); use pmu_fcomb_f32", F2_MAXMID + 2, nl);
  PMU_CHECK_ARG(aligned16(feat), "pmu_fcomb_softmax_accum_bf16: feat must be 16-byte aligned");
  PMU_CHECK_SUPPORTED((int64_t)B * HW < (1ll << 31), "pmu_fcomb_softmax_accum_bf16: B * HW must be below 2^31 (split the batch)");
  int cc_major = 0, dev = 0;
  PMU_CUDA(cudaGetDevice(&dev));
  PMU_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  PMU_CHECK_SUPPORTED(cc_major == 10, "pmu_fcomb_softmax_accum_bf16: needs an sm_100 device; found cc %d.x", cc_major);
  auto fn = f2_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PMU_ERR_CUDA; }
  CUtensorMap tmF;
  cuuint64_t dims[2] = {(cuuint64_t)F2_F, (cuuint64_t)((int64_t)B * HW)};
  cuuint64_t strides[1] = {(cuuint64_t)F2_F * 2};
  cuuint32_t box[2] = {(cuuint32_t)F2_F, 128};
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
  auto launch = [&](auto kern) -> int {
    PMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_TOTAL));
    kern<<<grid, F2_THREADS, F2_SMEM_TOTAL, (cudaStream_t)stream>>>(tmF, p, mu, sigma, eps, w0, b0, wmid, bmid, wlast, blast, slice_sums);
    PMU_LAUNCH_CHECK();
    return PMU_OK;
  };
  if (C <= 4) return f16 ? launch(fcomb_ts_kernel<4, true>) : launch(fcomb_ts_kernel<4, false>);
  return f16 ? launch(fcomb_ts_kernel<8, true>) : launch(fcomb_ts_kernel<8, false>);
}
