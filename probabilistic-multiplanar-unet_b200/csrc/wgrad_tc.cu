// Convolution WEIGHT GRADIENT on the 5th-gen tensor cores (tcgen05 + TMEM, TMA-fed), bf16 NHWC
// operands, fp32 accumulation — the wgrad half of nn.Conv2d's backward (unet_parts.py:15,18;
// probabilistic_unet.py:38,43) for the bf16 training mode (BASELINE config 4).
//
//   dW[co][tap][ci] = sum_{b,h,w} dY[b,h,w,co] * X[b,h+dy,w+dx,ci]        tap = (dy+1)*3 + (dx+1)
//
// GEMM view: the reduction runs over PIXELS, so both operands are "MN-major" for the UMMA — for
// one k (= one pixel) the 64 channels are contiguous — which is exactly how a TMA box
// {64 ch, TW, TH, TB} of the NHWC tensor lands in shared memory (rows of 128 B, 128-byte swizzle):
//   A (M = 128) = two "units" of 64 input channels, a unit being (tap, 64-channel block): the box of X
//                 shifted by the tap, out-of-image pixels zero-filled by TMA (= the conv padding).  With
//                 Cin = 64 the two units of an M tile are two different TAPS of the same channels.
//   B (N <= 256) = N/64 boxes of dY (unshifted).
//   K = 128 pixels per stage = 8 UMMAs of K = 16 (descriptor start += 16 rows * 128 B).
// The K range (all pixel tiles of the batch) is split across CTAs; every CTA adds its fp32 partial
// [128 x N] from TMEM into dW with coalesced atomics (a warp = 32 consecutive ci of one (co, tap)).
#include <cudaTypedefs.h>

#include "pmu_common.cuh"
#include "sm100_ptx.cuh"
#include "ctx.cuh"

namespace pmu {

using namespace ptx;

constexpr int WT_THREADS = 192;     // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
constexpr int WT_BOX = 128 * 128;   // one box: 128 pixels x 64 channels bf16 = 16 KB

struct WgradTcParams {
  int B, H, W, C0, C1, Cout, ntaps;
  int TW, TH, TB, tiles_w, tiles_h, tiles_b;
  int units;          // ntaps * Cin / 64
  int pairs;          // ceil(units / 2)
  int n_blocks;       // Cout / N
  int ksplit, kblocks_per_split, kblocks;
  int store;          // 1: the tile is not split and dw is overwritten: plain stores instead of atomics
};

// MN-major operand, 128-byte swizzle: 64 MN elements (128 B) contiguous, k rows 128 B apart, 8-row groups
// SBO = 1024 B apart, the next 64-element MN block LBO bytes away (cute: ((8,n),(8,k)):((1,LBO),(8,SBO))).
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// bf16 x bf16 -> fp32 with BOTH operands MN-major (bits 15 / 16 = transpose A / B)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

template <int N, int STAGES>
struct WgradSmem {
  static constexpr int STAGE_BYTES = (2 + N / 64) * WT_BOX;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;          // full[S], empty[S], done
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
  static constexpr int DYN_BYTES = TMEM_PTR_OFF + 16;
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget");
};

template <int N, int STAGES>
__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                const __grid_constant__ CUtensorMap tmDY, const WgradTcParams p, float* __restrict__ dw) {
  using L = WgradSmem<N, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t bar_full = smem_base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_done = bar_empty + STAGES * 8;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem_raw + L::TMEM_PTR_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cin = p.C0 + p.C1;
  const int cblocks = Cin / 64;

  // work item: (pair of A units, N block, K split)
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int nb = item % p.n_blocks;
  const int pi = item / p.n_blocks;
  const int u0 = 2 * pi, u1 = (2 * pi + 1 < p.units) ? 2 * pi + 1 : 2 * pi;   // a missing second unit repeats the first
  const int kb_lo = ks * p.kblocks_per_split;
  const int kb_hi = min(p.kblocks, kb_lo + p.kblocks_per_split);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmX0);
    if (p.C1 > 0) prefetch_tensormap(&tmX1);
    prefetch_tensormap(&tmDY);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  constexpr int TCOLS = (N <= 32) ? 32 : (N <= 64) ? 64 : (N <= 128) ? 128 : 256;
  if (warp == 1) tmem_alloc<TCOLS>(smem_base + L::TMEM_PTR_OFF);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {      // single provably-active thread: descriptors stay in uniform registers
      uint32_t kc = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++kc) {
        int t = kb;
        const int w0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
        const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
        const int b0 = t * p.TB;
        const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1u;
        mbar_wait(bar_empty + s * 8, ph ^ 1u);
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        mbar_arrive_expect_tx(bar_full + s * 8, L::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int u = j ? u1 : u0;
          const int tap = u / cblocks, c = (u % cblocks) * 64;
          const int dy = (p.ntaps == 9) ? tap / 3 - 1 : 0, dx = (p.ntaps == 9) ? tap % 3 - 1 : 0;
          if (c < p.C0) tma_load_4d(sa + j * WT_BOX, &tmX0, bar_full + s * 8, c, w0 + dx, h0 + dy, b0);
          else          tma_load_4d(sa + j * WT_BOX, &tmX1, bar_full + s * 8, c - p.C0, w0 + dx, h0 + dy, b0);
        }
#pragma unroll
        for (int j = 0; j < N / 64; ++j)
          tma_load_4d(sa + (2 + j) * WT_BOX, &tmDY, bar_full + s * 8, nb * N + j * 64, w0, h0, b0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {      // single provably-active thread: descriptors stay in uniform registers
      constexpr uint32_t idesc = umma_idesc_bf16_mn(128, N);
      uint32_t kc = 0;
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++kc) {
        const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1u;
        mbar_wait(bar_full + s * 8, ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        const uint64_t adesc = umma_smem_desc_mn_sw128(sa, WT_BOX);
        const uint64_t bdesc = umma_smem_desc_mn_sw128(sa + 2 * WT_BOX, WT_BOX);
#pragma unroll
        for (int k = 0; k < 8; ++k)      // 16 pixels (= 16 rows of 128 B = 2048 B) per UMMA
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (uint32_t)((kc | k) != 0));
        umma_commit(bar_empty + s * 8);
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // =========================== epilogue: TMEM -> atomics into dW[co][tap][ci] ===========================
    const int q = warp & 3;
    const int r = q * 32 + lane;                      // accumulator row = (unit r / 64, channel r % 64)
    const int u = (r < 64) ? u0 : u1;
    const bool valid = (r < 64) || (2 * pi + 1 < p.units);
    const int tap = u / cblocks, ci = (u % cblocks) * 64 + (r & 63);
    if (kb_hi > kb_lo) {
      mbar_wait(bar_done, 0);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(trow + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
          if (p.store) {
            // this CTA holds the whole pixel sum of its tile and dw is write-only: plain coalesced stores
            // (128 B per warp and column) instead of 128 x N atomics
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = nb * N + c0 + j;
              dw[((int64_t)co * p.ntaps + tap) * Cin + ci] = __uint_as_float(v[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = nb * N + c0 + j;
              atomicAdd(dw + ((int64_t)co * p.ntaps + tap) * Cin + ci, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tmem_base);
}

// operand descriptor through the launch context (ctx.cuh): cached when a context is bound
static int wt_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int TW, int TH, int TB) {
  TensorMapSpec s{};
  s.ptr = ptr; s.rank = 4; s.dtype = (int)CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; s.swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  s.l2promo = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B; s.oob = (int)CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE;
  s.dims[0] = (uint64_t)C; s.dims[1] = (uint64_t)W; s.dims[2] = (uint64_t)H; s.dims[3] = (uint64_t)B;
  s.strides[0] = (uint64_t)C * 2; s.strides[1] = (uint64_t)W * C * 2; s.strides[2] = (uint64_t)H * W * C * 2;
  s.box[0] = 64; s.box[1] = (uint32_t)TW; s.box[2] = (uint32_t)TH; s.box[3] = (uint32_t)TB;
  return tensor_map(m, s, "wgrad operand");
}
static int wt_pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <int N, int STAGES>
static int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& dy, const WgradTcParams& p,
                        float* dw, int64_t grid, cudaStream_t st) {
  using L = WgradSmem<N, STAGES>;
  auto kern = wgrad_tc_kernel<N, STAGES>;
  { const int rc_ = set_max_dyn_smem(reinterpret_cast<const void*>(kern), L::DYN_BYTES); if (rc_) return rc_; }
  kern<<<(unsigned)grid, WT_THREADS, L::DYN_BYTES, st>>>(x0, x1, dy, p, dw);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

}  // namespace pmu

using namespace pmu;

extern "C" int pmu_conv_wgrad_bf16(const void* x0, int C0, const void* x1, int C1, const void* dy, float* dw, int B,
                                   int H, int W, int Cout, int ntaps, int overwrite, void* stream) {
  PMU_CHECK_ARG(x0 && dy && dw && (C1 == 0 || x1), "pmu_conv_wgrad_bf16: null pointer");
  PMU_CHECK_ARG(ntaps == 9 || ntaps == 1, "pmu_conv_wgrad_bf16: ntaps must be 9 or 1 (got %d)", ntaps);
  PMU_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cout > 0 && C0 > 0 && C1 >= 0, "pmu_conv_wgrad_bf16: bad shape");
  PMU_CHECK_SUPPORTED(C0 % 64 == 0 && C1 % 64 == 0 && Cout % 64 == 0,
                      "pmu_conv_wgrad_bf16: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  PMU_CHECK_ARG(aligned16(x0) && aligned16(dy) && (!x1 || aligned16(x1)), "pmu_conv_wgrad_bf16: pointers must be 16-byte aligned");
  int cc_major = 0;
  { const int rc_ = device_cc_major(&cc_major); if (rc_) return rc_; }
  PMU_CHECK_SUPPORTED(cc_major == 10, "pmu_conv_wgrad_bf16: needs an sm_100 device (tcgen05/TMEM); found cc %d.x", cc_major);

  WgradTcParams p;
  p.B = B; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.Cout = Cout; p.ntaps = ntaps;
  p.TW = std::min(16, wt_pow2ceil(W));
  p.TH = std::min(128 / p.TW, wt_pow2ceil(H));
  p.TB = 128 / (p.TW * p.TH);
  p.tiles_w = cdiv(W, p.TW); p.tiles_h = cdiv(H, p.TH); p.tiles_b = cdiv(B, p.TB);
  p.kblocks = p.tiles_w * p.tiles_h * p.tiles_b;
  const int Cin = C0 + C1;
  p.units = ntaps * (Cin / 64);
  p.pairs = (p.units + 1) / 2;
  const int N = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  p.n_blocks = Cout / N;
  const int64_t tiles = (int64_t)p.pairs * p.n_blocks;
  // Split of the pixel range across CTAs.  A CTA costs its k-blocks (8 UMMAs of N columns each) plus its epilogue, and the
  // epilogue of a split tile — 128 x N fp32 atomics — costs more than eight k-blocks (measured on the 1024-channel layers
  // of a batch-8 step: 576 CTAs of 8 k-blocks took 55 us, i.e. ~10 us of atomics against 4.4 us of UMMAs per CTA), while an
  // unsplit tile leaves with plain read-modify-write stores.  Pick the split that minimises waves x (k-block time + epilogue)
  // (overwrite = 0, the accumulating form, always leaves with atomics).
  int ksplit = 1;
  {
    // a k-block: 8 UMMAs of N columns (0.55 us at N = 256) or, for narrow N, the 2 + N / 64 operand boxes it loads
    const double kb_us = std::max(0.55 * N / 256.0, 0.33), epi_store_us = 3.0 * N / 256.0, epi_atomic_us = 10.0 * N / 256.0;
    const int64_t slots = (int64_t)sm_count();                                 // one resident CTA per SM (192 KB of stages)
    double best = 1e30;
    const int kmax = (int)std::min<int64_t>(p.kblocks, 256);
    for (int ks = 1; ks <= kmax; ++ks) {
      const int kper = cdiv(p.kblocks, ks);
      if (ks > 1 && cdiv(p.kblocks, kper) != ks) continue;                      // same split as a smaller ks
      const double waves = (double)cdiv64(tiles * ks, slots);
      const double t = waves * (kper * kb_us + ((ks == 1 && overwrite) ? epi_store_us : epi_atomic_us));
      if (t < best - 1e-9) { best = t; ksplit = ks; }
    }
  }
  p.kblocks_per_split = cdiv(p.kblocks, ksplit);
  p.ksplit = cdiv(p.kblocks, p.kblocks_per_split);
  p.store = (overwrite && p.ksplit == 1) ? 1 : 0;
  if (overwrite && !p.store)      // split tiles add their partials: the library clears the accumulator itself
    PMU_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * ntaps * (C0 + C1), (cudaStream_t)stream));
  const int64_t grid = tiles * p.ksplit;
  PMU_CHECK_ARG(grid > 0 && grid < (1ll << 31), "pmu_conv_wgrad_bf16: grid too large");

  CUtensorMap mx0, mx1, mdy;
  int rc = wt_act_map(&mx0, x0, B, H, W, C0, p.TW, p.TH, p.TB);
  if (rc) return rc;
  if (C1 > 0) { rc = wt_act_map(&mx1, x1, B, H, W, C1, p.TW, p.TH, p.TB); if (rc) return rc; }
  else mx1 = mx0;
  rc = wt_act_map(&mdy, dy, B, H, W, Cout, p.TW, p.TH, p.TB);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 256) return launch_wgrad<256, 2>(mx0, mx1, mdy, p, dw, grid, st);
  if (N == 128) return launch_wgrad<128, 3>(mx0, mx1, mdy, p, dw, grid, st);
  return launch_wgrad<64, 4>(mx0, mx1, mdy, p, dw, grid, st);
}
