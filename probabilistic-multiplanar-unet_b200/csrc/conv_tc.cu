// K2 — implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), TMA-fed.
//
// Replaces nn.Conv2d(3x3,pad 1)+BatchNorm2d(eval, folded)+ReLU (unet_parts.py:15-20,
// probabilistic_unet.py:38-45), the F.pad/torch.cat of Up.forward (unet_parts.py:58-66, as a
// two-source K loop — the concat is never materialised) and nn.ConvTranspose2d(k2,s2)
// (unet_parts.py:52, as a 1-tap GEMM with N = 4*Cout and a pixel-shuffle store).
//
// GEMM view:  D[M = 128 pixels][N = BN couts] += A[M][K] * B[N][K]^T,  K = taps x Cin.
//   A tile : 128 output pixels = a (TB x TH x TW) brick of the NHWC bf16 activation tensor,
//            64 channels wide.  One TMA 4-D box load per (tap, 64-channel chunk); the tap
//            shift (dy,dx) is applied to the box coordinates and TMA's out-of-bounds zero
//            fill implements the conv padding — no im2col buffer, no halo bookkeeping.
//   B tile : BN rows of the packed weight matrix [Ntot][taps*Cin] (K-major), one TMA 2-D box.
//   Both land in shared memory in the 128-byte-swizzled K-major layout UMMA reads directly.
//   D      : fp32 accumulator in TMEM (BN columns x 128 lanes), read back with tcgen05.ld.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2..5 = epilogue (bias + ReLU + bf16 pack + NHWC store).
// Pipeline: STAGES-deep smem ring with full/empty mbarriers; tcgen05.commit releases slots.
// Persistent CTAs (one per SM) with a double-buffered TMEM accumulator: the epilogue of tile i
// overlaps the main loop of tile i+1.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "pmu_common.cuh"
#include "h16.cuh"
#include "sm100_ptx.cuh"
#include "ctx.cuh"

namespace pmu {

using namespace ptx;

constexpr int TC_BM = 128;       // pixels per tile (UMMA M)
constexpr int TC_BK = 64;        // channels per k-block (128 B of bf16 = one swizzle row)
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;

struct ConvTcParams {
  int B, H, W;        // input extents
  int C0, C1;         // channels of the two K-loop sources (multiples of 64; C1 may be 0)
  int Cout;           // output channels (per phase for convT)
  int ntaps;          // 9 = conv3x3, 1 = conv1x1, 4 = convT2x2 (N = 4*Cout, one K tap)
  int relu;
  int TW, TH, TB;     // tile brick, TW*TH*TB == 128
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;        // Ntot / BN
  int pool_mode;      // -1 none; PMU_POOL_MAX / PMU_POOL_AVG_CEIL: also emit the 2x2-pooled map
  int tma_store;      // a full-resolution output exists: it leaves through the smem staging tile + TMA tensor stores
  int f16;            // the 16-bit operands and outputs are IEEE f16 (inference) instead of bf16 (training), see h16.cuh
  double* stats = nullptr;  // training: per-channel {sum, sum^2} of the stored output y, stats[2 * co + k] += (BatchNorm batch
                            // statistics fused into the epilogue, unet_parts.py:16 / probabilistic_unet.py:39 in train() mode)
  int bias_bstride = 0;  // > 0: the bias is per image, bias[b * bias_bstride + co] (Fcomb's first layer after the split of
                         // its latent part, probabilistic_unet.py:167-176); needs TB == 1 (a tile lies in one image)
};

// RESW > 0 (the Cout = 64 transposed convolution, see convt_pair_epilogue): the layer's whole weight matrix — RESW k-blocks of [BN][64] — is loaded
// once per CTA and stays resident; only the activation boxes stream through the ring.  For a single-N-tile layer with a
// short K (the 128 -> 64 transposed convolution: N = 4 * 64 = 256, K = 128) the generic kernel re-fetches 64 KB of
// weights from L2 for every 128-pixel tile, as much as the tile writes.
template <int BN, int STAGES, int NSTG = 1, int RESW = 0>
struct ConvTcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int WRES_BYTES = RESW * B_BYTES;
  static constexpr int RING_OFF = WRES_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + (RESW ? 0 : B_BYTES);
  static constexpr int STG_OFF = RING_OFF + STAGES * STAGE_BYTES; // output staging tile [128 px][64 ch] bf16, 128B swizzle
  static constexpr int STG_BYTES = TC_BM * 128;                    // x NSTG buffers
  static constexpr int BAR_OFF = STG_OFF + NSTG * STG_BYTES;             // full[S], empty[S], tmem_full[2], tmem_empty[2][, wfull]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 4 + (RESW ? 1 : 0)) * 8;
  static constexpr int BIAS_OFF = ((TMEM_PTR_OFF + 4 + 15) / 16) * 16;   // BN floats
  static constexpr int TOTAL = BIAS_OFF + BN * 4;
  static constexpr int DYN_BYTES = TOTAL;            // base is 1024 B aligned (__align__ + runtime check)
};

// ------------------------------------------------------------------------------------
// Epilogue shared by the convolution kernels (warps 2..5, 128 threads, thread = tile row):
// TMEM -> bias + ReLU -> bf16 -> swizzled smem staging tile -> TMA tensor store, plus the fused
// 2x2 pooling.  The tile brick is p.TW x p.TH pixels (16 x 8 or 8 x 16): a warp always holds
// complete pooling windows in lanes {l, l^1, l^TW}.
// ------------------------------------------------------------------------------------
// bias + [ReLU] + 16-bit pack of 32 accumulator columns: one packed fp32 add (add.f32x2) + one cvt per output pair, the
// ReLU rides in the cvt.  (The scalar form — 2 FADD, 2 FMNMX, 1 F2FP and 2 LDS per pair — made this epilogue, not the
// MMA, the critical path of the 64/128-cout layers: ~3000 cycles per tile against 1152 of UMMA at 64 -> 64.)
__device__ __forceinline__ float4 lds128_f4(uint32_t addr);
template <bool RELU, bool F16>
__device__ __forceinline__ void bias_pack32(const uint32_t (&r)[32], uint32_t bs_addr, uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = lds128_f4(bs_addr + j * 16);
    pk[2 * j] = add_pack16<RELU, F16>(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), b4.x, b4.y);
    pk[2 * j + 1] = add_pack16<RELU, F16>(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]), b4.z, b4.w);
  }
}
__device__ __forceinline__ void bias_pack32(const uint32_t (&r)[32], uint32_t bs_addr, uint32_t (&pk)[16], int relu, int f16) {
  if (f16) { if (relu) bias_pack32<true, true>(r, bs_addr, pk); else bias_pack32<false, true>(r, bs_addr, pk); }
  else { if (relu) bias_pack32<true, false>(r, bs_addr, pk); else bias_pack32<false, false>(r, bs_addr, pk); }
}
__device__ __forceinline__ float4 lds128_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ uint32_t lds32_u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32_f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct EpiCtx {
  uint32_t smem_base, stg_off, bar_tfull, bar_tempty, tmem_base;
  float* bias_s;
};

template <int BN, int NSTG, bool STATS = false>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, const EpiCtx& e, const CUtensorMap* tmY0p,
                                              const CUtensorMap* tmY1p, const CUtensorMap* tmY2p,
                                              const CUtensorMap* tmY3p, const float* __restrict__ bias,
                                              __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y_pool,
                                              int total_tiles, int warp, int lane) {
  const uint32_t smem_base = e.smem_base, bar_tfull = e.bar_tfull, bar_tempty = e.bar_tempty, tmem_base = e.tmem_base;
  float* bias_s = e.bias_s;
  constexpr int STG_BYTES = TC_BM * 128;
  const int et = threadIdx.x - 64;  // 0..127
  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int m = q * 32 + lane;       // row of the tile == pixel index in the brick
  const int tx = m % p.TW, ty = (m / p.TW) % p.TH, tb = m / (p.TW * p.TH);
  uint32_t iter = 0, stg_count = 0;
  // ---- fused BatchNorm statistics (p.stats, training): every thread keeps {sum, sum^2} of ONE channel pair over the 32
  // pixel rows of its warp, per 64-channel group, across all tiles of this CTA (they share the N tile unless the grid is
  // not a multiple of n_tiles); read back from the bf16 staging tile — the values the BatchNorm apply pass will read from
  // HBM — while the TMA store of the same tile is in flight.  Flushed once: the four warps' partials meet in the (idle)
  // staging buffer and leave as 2 * BN fp64 atomics per CTA.
  float bst[STATS ? BN / 64 : 1][4];
#pragma unroll
  for (int g = 0; g < (STATS ? BN / 64 : 1); ++g) bst[g][0] = bst[g][1] = bst[g][2] = bst[g][3] = 0.f;
  int stat_n0 = -1;
  const int cp = lane;                       // channel pair of the 64-channel group this thread accumulates
  auto stats_flush = [&]() {
    if (et == 0) tma_store_wait_read0();     // every staging buffer is idle
    named_bar_sync(2, 128);
    const uint32_t scr = smem_base + e.stg_off;
#pragma unroll
    for (int g = 0; g < BN / 64; ++g)
#pragma unroll
      for (int k = 0; k < 4; ++k) { sts32_f(scr + (uint32_t)(((g * 4 + q) * 128 + cp * 4 + k) * 4), bst[g][k]); bst[g][k] = 0.f; }
    named_bar_sync(3, 128);
    for (int o = et; o < 2 * BN; o += 128) {
      const int c = o >> 1, k2 = o & 1, g = c >> 6, idx = ((c & 63) >> 1) * 4 + k2 * 2 + (c & 1);
      float sum = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) sum += lds32_f(scr + (uint32_t)(((g * 4 + qq) * 128 + idx) * 4));
      const int co = stat_n0 + c;
      if (co < p.Cout) atomicAdd(p.stats + 2 * (int64_t)co + k2, (double)sum);
    }
  };
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
    const uint32_t as = iter & 1u, aph = (iter >> 1) & 1u;
    const int n_tile = tile % p.n_tiles;
    int m_tile = tile / p.n_tiles;
    const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
    const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
    const int b0 = m_tile * p.TB;
    const int n0 = n_tile * BN;
    const int co_base = (p.ntaps == 4) ? 0 : n0;   // convT: output channel / phase are resolved per 64-column group
    float* bs = bias_s;
    named_bar_sync(4, 128);          // every warp is done reading the previous tile's bias
    for (int i = et; i < BN; i += 128)
      bs[i] = bias ? __ldg(bias + (int64_t)b0 * p.bias_bstride + ((p.ntaps == 4) ? (n0 + i) % p.Cout : n0 + i)) : 0.f;
    named_bar_sync(1, 128);          // bias visible

    const int b = b0 + tb, h = h0 + ty, w = w0 + tx;
    const bool valid = (b < p.B) && (h < p.H) && (w < p.W);
    uint32_t vmask = 0;
    if constexpr (STATS) {
      if (stat_n0 >= 0 && stat_n0 != n0) stats_flush();
      stat_n0 = n0;
      vmask = __ballot_sync(0xffffffffu, valid);      // bit i: row q * 32 + i of the tile is a pixel of the image
    }
    // fused 2x2 pooling (a warp holds 32 / TW complete image rows of the brick, so every pooling
    // window lives in lanes {l, l^1, l^TW, l^(1|TW)} of one warp — no extra pass over HBM)
    __nv_bfloat16* dstp = nullptr;
    if (p.pool_mode >= 0)
      dstp = y_pool + (((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * p.Cout + co_base;

    mbar_wait(bar_tfull + as * 8, aph);
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int g0 = 0; g0 < BN; g0 += 64) {
      const uint32_t stg = smem_base + e.stg_off + (NSTG > 1 ? (stg_count % NSTG) * STG_BYTES : 0);
      const uint32_t stg_row = stg + m * 128;
      if (p.tma_store) {
        // the TMA store that used this staging buffer must have finished READING it before it is rewritten
        if (et == 0) { if (NSTG > 1) tma_store_wait_read1(); else tma_store_wait_read0(); }
        named_bar_sync(2, 128);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = g0 + half * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_d + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {
          // all of this thread's TMEM reads of the stage are done: hand it back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(bar_tempty + as * 8);
        }
        uint32_t pk[16];
        bias_pack32(r, smem_u32(bs) + (uint32_t)c0 * 4, pk, p.relu, p.f16);
        if (p.tma_store) {
          // staging row m, 16-byte chunk (half*4 + j), 128-byte swizzle (chunk ^ (row & 7))
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128_u32(stg_row + ((((half * 4 + j) ^ (m & 7)) & 7) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        if (p.pool_mode >= 0) {
          // Halving exchange over the window lanes {l, l^1, l^TW, l^(1|TW)}.  (A butterfly that leaves the whole pooled
          // row in all four lanes and lets one of them write 64 B costs 32 shuffles + 32 max per 32 channels and thread:
          // 80-95 us per launch on the 64-cout layers, profiles/r02_experiments.txt.)  Here a lane keeps half of its
          // channels per step and sends the other half: step 1 (lane ^ 1) 8 registers, step 2
          // (lane ^ TW) 4 registers (8 fp32 sums for the average) — 12 (16) shuffles — and every lane ends up with 8
          // channels of the pooled pixel, which it writes itself (16 B each; the four lanes' pieces are contiguous).
          // max is exact; the average adds (a + b) + (c + d) in fp32 and rounds once.
          const bool odd = (lane & 1) != 0, up = (lane & p.TW) != 0;
          uint32_t o4[4];
          if (p.pool_mode == PMU_POOL_MAX) {
            uint32_t k8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t send = odd ? pk[j] : pk[j + 8], keep = odd ? pk[j + 8] : pk[j];
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
              k8[j] = p.f16 ? max16x2<true>(keep, recv) : max16x2<false>(keep, recv);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t send = up ? k8[j] : k8[j + 4], keep = up ? k8[j + 4] : k8[j];
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, p.TW);
              o4[j] = p.f16 ? max16x2<true>(keep, recv) : max16x2<false>(keep, recv);
            }
          } else {
            float lo8[8], hi8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t send = odd ? pk[j] : pk[j + 8], keep = odd ? pk[j + 8] : pk[j];
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
              const float2 a = p.f16 ? unpack16<true>(keep) : unpack16<false>(keep), c = p.f16 ? unpack16<true>(recv) : unpack16<false>(recv);
              lo8[j] = a.x + c.x;
              hi8[j] = a.y + c.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float slo = up ? lo8[j] : lo8[j + 4], shi = up ? hi8[j] : hi8[j + 4];
              const float klo = up ? lo8[j + 4] : lo8[j], khi = up ? hi8[j + 4] : hi8[j];
              const float lo = klo + __shfl_xor_sync(0xffffffffu, slo, p.TW), hi = khi + __shfl_xor_sync(0xffffffffu, shi, p.TW);
              o4[j] = p.f16 ? pack16_rn<true>(lo * 0.25f, hi * 0.25f) : pack16_rn<false>(lo * 0.25f, hi * 0.25f);
            }
          }
          if (valid)    // a window never straddles the image edge (even H, W; bricks start at even coordinates)
            *reinterpret_cast<uint4*>(dstp + c0 + (odd ? 16 : 0) + (up ? 8 : 0)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
      if (p.tma_store) {
        fence_proxy_async_smem();      // staging writes (generic proxy) -> visible to the TMA engine
        named_bar_sync(3, 128);
        if (et == 0) {
          // one tensor store per 64-channel group: full 128-byte lines, out-of-bounds pixels clipped by TMA
          if (p.ntaps == 4) {
            const int n = n0 + g0, ij = n / p.Cout, co = n - ij * p.Cout;
            const CUtensorMap* tm = (ij == 0) ? tmY0p : (ij == 1) ? tmY1p : (ij == 2) ? tmY2p : tmY3p;
            tma_store_4d(tm, stg, co, w0, h0, b0);
          } else {
            tma_store_4d(tmY0p, stg, n0 + g0, w0, h0, b0);
          }
          tma_store_commit();
        }
        if constexpr (STATS) {
          // column pair cp of the staged tile over this warp's 32 rows (conflict-free: a warp reads one 128-byte row per step)
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint32_t col = (uint32_t)(cp & 3) * 4u, c16 = (uint32_t)cp >> 2;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const uint32_t r = (uint32_t)(q * 32 + i);
            uint32_t v = lds32_u(stg + r * 128u + (((c16 ^ (r & 7u)) & 7u) << 4) + col);
            if (!((vmask >> i) & 1u)) v = 0u;
            const float2 f = p.f16 ? unpack16<true>(v) : unpack16<false>(v);
            s0 += f.x; s1 += f.y; q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
          }
          const int g = g0 >> 6;
          bst[g][0] += s0; bst[g][1] += s1; bst[g][2] += q0; bst[g][3] += q1;
        }
        ++stg_count;
      }
    }
  }
  if constexpr (STATS) { if (stat_n0 >= 0) stats_flush(); }
  if (p.tma_store && et == 0) tma_store_wait_all();
}

// Epilogue of the one-N-tile transposed convolution (Cout = 64, N = 4 * 64).
// The shared epilogue stores every phase (i, j) through its own strided tensor map: 128 scattered 128-byte pieces per
// store (pixel stride 256 B) — measured, that layer ran at ~6400 cycles per tile with nothing saturated (L2 28 %,
// DRAM 44 %, tensor pipe 16 %): it waited for its stores (0.221 -> 0.148 ms per batch of 64 with this epilogue).  Here the two column parities j = 0, 1 of one output-row parity i
// are staged INTERLEAVED, staging row = (tb, ty, 2 * tx + j), so that one tensor store per i writes 2 * TW * 128 B = 4 KB
// contiguous runs of the output row 2 * h + i: half as many stores, each over whole rows.
template <int NSTG>
__device__ __forceinline__ void convt_pair_epilogue(const ConvTcParams& p, const EpiCtx& e, const CUtensorMap* tmP0,
                                                    const CUtensorMap* tmP1, const float* __restrict__ bias,
                                                    int total_tiles, int warp, int lane) {
  constexpr int BN = 256, PAIR_BYTES = 2 * TC_BM * 128, NPAIR = NSTG / 2;
  static_assert(NSTG >= 2 && NSTG % 2 == 0, "pair staging buffers are two 16 KB tiles each");
  const uint32_t smem_base = e.smem_base, bar_tfull = e.bar_tfull, bar_tempty = e.bar_tempty, tmem_base = e.tmem_base;
  float* bs = e.bias_s;
  const int et = threadIdx.x - 64;   // 0..127
  const int q = warp & 3;
  const int m = q * 32 + lane;
  const int tx = m % p.TW, ty = (m / p.TW) % p.TH, tb = m / (p.TW * p.TH);
  uint32_t iter = 0, pair_count = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
    const uint32_t as = iter & 1u, aph = (iter >> 1) & 1u;
    int m_tile = tile;                               // one N tile
    const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
    const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
    const int b0 = m_tile * p.TB;
    named_bar_sync(4, 128);          // every warp is done reading the previous tile's bias
    for (int i = et; i < BN; i += 128) bs[i] = bias ? __ldg(bias + i % p.Cout) : 0.f;
    named_bar_sync(1, 128);          // bias visible
    mbar_wait(bar_tfull + as * 8, aph);
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int i2 = 0; i2 < 2; ++i2) {                 // output row parity
      const uint32_t stg = smem_base + e.stg_off + (pair_count % NPAIR) * PAIR_BYTES;
      // the store that used this pair buffer must have finished READING it before it is rewritten
      if (et == 0) { if (NPAIR > 1) tma_store_wait_read1(); else tma_store_wait_read0(); }
      named_bar_sync(2, 128);
#pragma unroll
      for (int j = 0; j < 2; ++j) {                  // output column parity: GEMM columns [(2 i + j) * 64, +64)
        const int row = (tb * p.TH + ty) * (2 * p.TW) + 2 * tx + j;
        const uint32_t stg_row = stg + (uint32_t)row * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = (i2 * 2 + j) * 64 + half * 32;
          uint32_t r[32];
          tmem_ld_32x32(tmem_d + (uint32_t)c0, r);
          tmem_ld_wait();
          if (c0 + 32 >= BN) {       // last TMEM read of this accumulator stage
            tcgen05_fence_before();
            mbar_arrive(bar_tempty + as * 8);
          }
          uint32_t pk[16];
          bias_pack32(r, smem_u32(bs) + (uint32_t)c0 * 4, pk, p.relu, p.f16);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sts128_u32(stg_row + ((((half * 4 + k) ^ (row & 7)) & 7) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        }
      }
      fence_proxy_async_smem();      // staging writes (generic proxy) -> visible to the TMA engine
      named_bar_sync(3, 128);
      if (et == 0) {
        // box {64 ch, 2 TW output pixels, TH rows of parity i2, TB}: rows 2 h + i2, pixels 2 w0 .. 2 w0 + 2 TW - 1
        tma_store_4d(i2 ? tmP1 : tmP0, stg, 0, 2 * w0, h0, b0);
        tma_store_commit();
      }
      ++pair_count;
    }
  }
  if (et == 0) tma_store_wait_all();
}

// Persistent kernel: one CTA per SM loops over output tiles (tile = blockIdx.x + i*gridDim.x,
// N-tile fastest so concurrently running CTAs share activation bricks in L2).  The smem ring
// keeps running across tiles (the producer prefetches the next tile's operands while the
// current one is still in the tensor pipe) and the accumulator is double-buffered in TMEM
// (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i+1 and the
// setup cost (barrier init, TMEM allocation, descriptor prefetch) is paid once per SM.
template <int BN, int STAGES, int MINB, int NSTG, int RESW = 0, bool TPAIR = false, bool STATS = false>
__global__ void __launch_bounds__(TC_THREADS, MINB)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY0,
               const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
               const __grid_constant__ CUtensorMap tmY3, const ConvTcParams p,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
               __nv_bfloat16* __restrict__ y_pool) {
  using L = ConvTcSmem<BN, STAGES, NSTG, RESW>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  uint8_t* smem_gen = smem_raw;
  if ((smem_base & 1023u) != 0) __trap();   // TMA / UMMA tiles need 1024 B alignment
  const uint32_t bar_full = smem_base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;   // [2] accumulator stage ready for the epilogue
  const uint32_t bar_tempty = bar_tfull + 2 * 8;       // [2] accumulator stage drained by the epilogue
  const uint32_t bar_wfull = bar_tempty + 2 * 8;       // RESW only: resident weights have landed
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem_gen + L::TMEM_PTR_OFF);
  float* bias_s = reinterpret_cast<float*>(smem_gen + L::BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cin = p.C0 + p.C1;
  const int kc_per_tap = Cin / TC_BK;
  const int k_taps = (p.ntaps == 9) ? 9 : 1;
  const int k_iters = k_taps * kc_per_tap;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA0);
    if (p.C1 > 0) prefetch_tensormap(&tmA1);
    prefetch_tensormap(&tmW);
    if (p.tma_store) prefetch_tensormap(&tmY0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + a * 8, 1);
      mbar_init(bar_tempty + a * 8, 128);   // every epilogue thread arrives
    }
    if (RESW) mbar_init(bar_wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(smem_base + L::TMEM_PTR_OFF);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      if (RESW) {   // the whole weight matrix, once (the launcher guarantees one N tile and k_iters <= RESW)
        mbar_arrive_expect_tx(bar_wfull, (uint32_t)k_iters * L::B_BYTES);
        for (int it = 0; it < k_iters; ++it) tma_load_2d(smem_base + it * L::B_BYTES, &tmW, bar_wfull, it * TC_BK, 0);
      }
      uint32_t kc = 0;   // k-block counter over the whole tile sequence
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
        const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
        const int b0 = m_tile * p.TB;
        const int n0 = n_tile * BN;
        for (int it = 0; it < k_iters; ++it, ++kc) {
          const uint32_t s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1u;
          mbar_wait(bar_empty + s * 8, ph ^ 1u);
          const int tap = it / kc_per_tap;
          const int c = (it - tap * kc_per_tap) * TC_BK;   // channel offset inside the concatenated K
          int dy = 0, dx = 0;
          if (p.ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
          const uint32_t sa = smem_base + L::RING_OFF + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(bar_full + s * 8, L::STAGE_BYTES);
          if (c < p.C0) tma_load_4d(sa, &tmA0, bar_full + s * 8, c, w0 + dx, h0 + dy, b0);
          else          tma_load_4d(sa, &tmA1, bar_full + s * 8, c - p.C0, w0 + dx, h0 + dy, b0);
          if (!RESW) tma_load_2d(sb, &tmW, bar_full + s * 8, tap * Cin + c, n0);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // one elected thread runs the whole issue loop: nothing but barrier polls between MMAs
    if (elect_one()) {
      const uint32_t idesc = p.f16 ? umma_idesc_f16(TC_BM, BN) : umma_idesc_bf16(TC_BM, BN);
      if (RESW) { mbar_wait(bar_wfull, 0); tcgen05_fence_after(); }
      uint32_t kc = 0, iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const uint32_t as = iter & 1u, aph = (iter >> 1) & 1u;
        mbar_wait(bar_tempty + as * 8, aph ^ 1u);     // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int it = 0; it < k_iters; ++it, ++kc) {
          const uint32_t s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1u;
          mbar_wait(bar_full + s * 8, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + L::RING_OFF + s * L::STAGE_BYTES;
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(RESW ? smem_base + it * L::B_BYTES : sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            // +32 B per UMMA_K inside the 128 B swizzle row: start-address field += 2
            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                      (uint32_t)((it | k) != 0));
          }
          umma_commit(bar_empty + s * 8);                        // frees the smem slot when these MMAs retire
          if (it == k_iters - 1) umma_commit(bar_tfull + as * 8);  // accumulator complete
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    EpiCtx ec{smem_base, (uint32_t)L::STG_OFF, bar_tfull, bar_tempty, tmem_base, bias_s};
    if constexpr (TPAIR) convt_pair_epilogue<NSTG>(p, ec, &tmY0, &tmY1, bias, total_tiles, warp, lane);
    else conv_epilogue<BN, NSTG, STATS>(p, ec, &tmY0, &tmY1, &tmY2, &tmY3, bias, y, y_pool, total_tiles, warp, lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------
// "Row-shift" conv3x3 for the small-N layers (Cout = 64 / 128), where the generic kernel is bound by
// operand traffic: one A box + one B box per (tap, chunk) is 24 KB per 128 MMA cycles at N = 64 —
// more bytes in flight than a CTA's shared memory can hold against the L2 latency.
// Here the tile is 16 rows x 8 px and an A box is the (16+2)-row x 8-px x 64-channel slab
//   {64 ch, 8, 18, 1} at x-shift dx.  Rows of the box are 128-byte swizzle rows ordered (h, w), so an
// 8-row swizzle group == one image row and the three dy taps of that dx are the SAME box read through
// descriptors offset by dy * 1024 B: 3 A loads per 64-channel chunk instead of 9 (A traffic / 2.7).
// A stage = one A box + the three weight taps (dy = 0..2) of that dx: 12 UMMAs per stage.
// RESB = number of resident 64-channel weight chunks (0 = weights stream with the A boxes).  RESB = 1 (Cin == 64,
// Cout == 64): all nine weight taps (72 KB) stay resident in smem for the whole kernel and only A streams (18 KB
// per 384 MMA cycles); the 64 -> 128 layers keep their nine 16 KB taps resident the same way (BN = 128, RESB = 1:
// 0.162 -> 0.148 ms per batch of 64).  (Keeping the 144 KB of the 128 -> 64 layer resident with a 3-stage A ring was
// measured slower, 0.652 -> 0.809 ms, and is gone: profiles/r02_experiments.txt.)
// ------------------------------------------------------------------------------------
template <int BN, int STAGES, int RESB>
struct ConvRsSmem {
  static constexpr int W_TAP = BN * 128;                           // one tap, one 64-channel chunk: [BN][64] bf16
  static constexpr int BOX_BYTES = 18 * 8 * 128;                   // 18,432
  static constexpr int WRES_BYTES = 9 * RESB * W_TAP;              // resident weights: [tap][chunk][BN][64]
  static constexpr int STAGE_BYTES = BOX_BYTES + (RESB ? 0 : 3 * W_TAP);
  static constexpr int RING_OFF = WRES_BYTES;
  static constexpr int STG_OFF = RING_OFF + STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STG_OFF + TC_BM * 128;            // full[S], empty[S], tfull[2], tempty[2], wfull
  static constexpr int TMEM_PTR_OFF = BAR_OFF + (2 * STAGES + 5) * 8;
  static constexpr int BIAS_OFF = ((TMEM_PTR_OFF + 4 + 15) / 16) * 16;
  static constexpr int DYN_BYTES = BIAS_OFF + BN * 4;
  static_assert(WRES_BYTES % 1024 == 0 && STAGE_BYTES % 1024 == 0, "tiles must stay 1024 B aligned");
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, int STAGES, int RESB, bool STATS = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY0, const ConvTcParams p,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y_pool) {
  using L = ConvRsSmem<BN, STAGES, RESB>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t bar_full = smem_base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 2 * 8;
  const uint32_t bar_wfull = bar_tempty + 2 * 8;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem_raw + L::TMEM_PTR_OFF);
  float* bias_s = reinterpret_cast<float*>(smem_raw + L::BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cin = p.C0 + p.C1;
  const int chunks = Cin / TC_BK;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA0);
    if (p.C1 > 0) prefetch_tensormap(&tmA1);
    prefetch_tensormap(&tmW);
    if (p.tma_store) prefetch_tensormap(&tmY0);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + a * 8, 1); mbar_init(bar_tempty + a * 8, 128); }
    mbar_init(bar_wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(smem_base + L::TMEM_PTR_OFF);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      if (RESB) {   // the nine weight taps (all RESB chunks of each), once
        mbar_arrive_expect_tx(bar_wfull, L::WRES_BYTES);
        for (int tap = 0; tap < 9; ++tap)
          for (int rc = 0; rc < RESB; ++rc)
            tma_load_2d(smem_base + (tap * RESB + rc) * L::W_TAP, &tmW, bar_wfull, tap * Cin + rc * TC_BK, 0);
      }
      uint32_t kc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
        const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
        const int b0 = m_tile;
        const int n0 = n_tile * BN;
        for (int ch = 0; ch < chunks; ++ch) {
          const int c = ch * TC_BK;
          for (int dx = 0; dx < 3; ++dx, ++kc) {
            const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1u;
            mbar_wait(bar_empty + s * 8, ph ^ 1u);
            const uint32_t sa = smem_base + L::RING_OFF + s * L::STAGE_BYTES;
            // rows h0-1 .. h0+16, pixels w0+dx-1 .. +7; out-of-image rows / pixels are zero-filled = conv padding
            mbar_arrive_expect_tx(bar_full + s * 8, L::STAGE_BYTES);
            if (c < p.C0) tma_load_4d(sa, &tmA0, bar_full + s * 8, c, w0 + dx - 1, h0 - 1, b0);
            else          tma_load_4d(sa, &tmA1, bar_full + s * 8, c - p.C0, w0 + dx - 1, h0 - 1, b0);
            if (!RESB) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
                tma_load_2d(sa + L::BOX_BYTES + dy * L::W_TAP, &tmW, bar_full + s * 8, (dy * 3 + dx) * Cin + c, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = p.f16 ? umma_idesc_f16(TC_BM, BN) : umma_idesc_bf16(TC_BM, BN);
      if (RESB) { mbar_wait(bar_wfull, 0); tcgen05_fence_after(); }
      uint32_t kc = 0, iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const uint32_t as = iter & 1u, aph = (iter >> 1) & 1u;
        mbar_wait(bar_tempty + as * 8, aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int ch = 0; ch < chunks; ++ch) {
          for (int dx = 0; dx < 3; ++dx, ++kc) {
            const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1u;
            mbar_wait(bar_full + s * 8, ph);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + L::RING_OFF + s * L::STAGE_BYTES;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              // tap (dy, dx): image rows dy .. dy+15 of the box = +dy swizzle groups of 8 rows (1024 B each)
              const uint64_t adesc = umma_smem_desc_sw128(sa + dy * 1024);
              const uint64_t bdesc = umma_smem_desc_sw128(RESB ? smem_base + (dy * 3 + dx) * L::W_TAP
                                                               : sa + L::BOX_BYTES + dy * L::W_TAP);
#pragma unroll
              for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                          (uint32_t)((ch | dx | dy | k) != 0));
            }
            umma_commit(bar_empty + s * 8);
            if (ch == chunks - 1 && dx == 2) umma_commit(bar_tfull + as * 8);
          }
        }
      }
    }
    __syncwarp();
  } else {
    EpiCtx ec{smem_base, (uint32_t)L::STG_OFF, bar_tfull, bar_tempty, tmem_base, bias_s};
    conv_epilogue<BN, 1, STATS>(p, ec, &tmY0, &tmY0, &tmY0, &tmY0, bias, y, y_pool, total_tiles, warp, lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem_base);
}


// ------------------------------------------------------------------------------------
// First layer (Cin = 1 -> 64) on the tensor cores.  The CUDA-core stencil (layers_bf16.cu) needs 576 FMAs per pixel
// and ran at 2.2 TB/s of output; here the 3x3 neighbourhood of every pixel is laid out as ONE K-major operand row
// (im2col in shared memory, built by four producer warps from a staged fp32 patch) and the 9-tap contraction is two
// K = 16 UMMAs per 128-pixel tile: k = 0..8 the bf16 high parts of the nine taps, k = 9..17 their bf16 low parts
// (x = hi + lo keeps the fp32 input's precision; the weight row repeats the nine weights for both), the rest zero.
// The kernel is then bound by its 128 B / pixel of output, written by the shared TMA-store epilogue.
// ------------------------------------------------------------------------------------
constexpr int CF_THREADS = 448;           // warp 1: MMA issuer, warps 2..5 / 6..9: epilogue groups 0 / 1, warps 10..13: im2col producers
struct ConvFirstSmem {
  static constexpr int A_BYTES = TC_BM * 128;                 // [128 px][64 k] bf16 rows of 128 B (k < 32 used), x2
  static constexpr int W_OFF = 2 * A_BYTES;                   // [64 couts][64 k]
  static constexpr int STG_OFF = W_OFF + 64 * 128;            // one staging tile per epilogue group
  static constexpr int PATCH_OFF = STG_OFF + 2 * TC_BM * 128; // fp32 (8+2) x (16+2) input patch, x2
  static constexpr int PATCH_FLOATS = 192;
  static constexpr int BAR_OFF = PATCH_OFF + 2 * PATCH_FLOATS * 4;   // a_full[2], a_empty[2], tfull[2], tempty[2]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + 8 * 8;
  static constexpr int BIAS_OFF = ((TMEM_PTR_OFF + 4 + 15) / 16) * 16;
  static constexpr int DYN_BYTES = BIAS_OFF + 64 * 4;
};

__global__ void __launch_bounds__(CF_THREADS, 2)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap tmY, const ConvTcParams p, const float* __restrict__ x,
                     const float* __restrict__ w, const float* __restrict__ bias) {
  using L = ConvFirstSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t bar_afull = smem_base + L::BAR_OFF, bar_aempty = bar_afull + 16, bar_tfull = bar_aempty + 16,
                 bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem_raw + L::TMEM_PTR_OFF);
  float* bias_s = reinterpret_cast<float*>(smem_raw + L::BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_b;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmY);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_afull + a * 8, 128); mbar_init(bar_aempty + a * 8, 1);
      mbar_init(bar_tfull + a * 8, 1); mbar_init(bar_tempty + a * 8, 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(smem_base + L::TMEM_PTR_OFF);
  // zero the operand tiles (k >= 18 must read as zero), then the weight tile: row co = [w(9) | w(9) | 0...]
  for (int i = threadIdx.x; i < (L::STG_OFF) / 16; i += CF_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 18; i += CF_THREADS) {
    const int co = i / 18, k = i % 18;
    const uint32_t off = (uint32_t)(co * 128 + ((((k >> 3) ^ (co & 7)) & 7) << 4) + (k & 7) * 2);
    const float wv = __ldg(w + co * 9 + (k % 9));
    *reinterpret_cast<uint16_t*>(smem_raw + L::W_OFF + off) = p.f16 ? cvt16<true>(wv) : cvt16<false>(wv);
  }
  for (int i = threadIdx.x; i < 64; i += CF_THREADS) bias_s[i] = bias ? __ldg(bias + i) : 0.f;
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = p.f16 ? umma_idesc_f16(TC_BM, 64) : umma_idesc_bf16(TC_BM, 64);
      const uint64_t bdesc = umma_smem_desc_sw128(smem_base + L::W_OFF);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait(bar_afull + s * 8, ph);
        mbar_wait(bar_tempty + s * 8, ph ^ 1u);
        tcgen05_fence_after();
        const uint64_t adesc = umma_smem_desc_sw128(smem_base + s * L::A_BYTES);
        umma_bf16(tmem_base + s * 64, adesc, bdesc, idesc, 0u);
        umma_bf16(tmem_base + s * 64, adesc + 2, bdesc + 2, idesc, 1u);
        umma_commit(bar_aempty + s * 8);
        umma_commit(bar_tfull + s * 8);
      }
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 10) {
    // =========================== epilogue: two groups of four warps, group e owns TMEM stage e ===========================
    // (the MMA is two instructions per tile, so the epilogue IS the critical path: tiles alternate between the groups)
    const int e = (warp - 2) >> 2;
    const int eg = threadIdx.x - 64 - e * 128;       // 0..127
    const int q = warp & 3;
    const int m = q * 32 + lane;                     // tile row == pixel of the 16 x 8 brick
    const uint32_t stg = smem_base + L::STG_OFF + e * (TC_BM * 128);
    const uint32_t stg_row = stg + m * 128;
    const uint32_t tmem_d = tmem_base + e * 64 + ((uint32_t)(q * 32) << 16);
    const uint32_t bs_addr = smem_u32(bias_s);
    uint32_t k = 0;
    for (int tile = blockIdx.x + e * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, ++k) {
      int m_tile = tile;
      const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
      const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
      const int b0 = m_tile;
      if (eg == 0) tma_store_wait_read0();           // the previous store of this group has finished reading the staging tile
      named_bar_sync(1 + 2 * e, 128);
      mbar_wait(bar_tfull + e * 8, k & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_d + (uint32_t)(half * 32), r);
        tmem_ld_wait();
        if (half == 1) { tcgen05_fence_before(); mbar_arrive(bar_tempty + e * 8); }
        uint32_t pk[16];
        bias_pack32(r, bs_addr + (uint32_t)(half * 32) * 4, pk, p.relu, p.f16);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128_u32(stg_row + ((((half * 4 + j) ^ (m & 7)) & 7) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(2 + 2 * e, 128);
      if (eg == 0) { tma_store_4d(&tmY, stg, 0, w0, h0, b0); tma_store_commit(); }
    }
    if (eg == 0) tma_store_wait_all();
  } else if (warp >= 10) {
    // =========================== im2col producers (128 threads, thread = tile row) ===========================
    const int m = threadIdx.x - 320;
    const int tx = m % p.TW, ty = m / p.TW;
    const int PW = p.TW + 2, PH = p.TH + 2;
    const int64_t HW = (int64_t)p.H * p.W;
    // the patch of tile i+1 is loaded into registers while tile i is built: the global-load latency (~1.5k cycles
    // per tile when exposed) was what bounded this kernel, not the epilogue
    auto load_patch = [&](int tile, float (&v)[2]) {
      int m_tile = tile;
      const int w0 = (m_tile % p.tiles_w) * p.TW; m_tile /= p.tiles_w;
      const int h0 = (m_tile % p.tiles_h) * p.TH; m_tile /= p.tiles_h;
      const int b0 = m_tile;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = m + 128 * j;
        const int r = e / PW, c = e % PW;
        const int gy = h0 - 1 + r, gx = w0 - 1 + c;
        v[j] = (e < PH * PW && tile < total_tiles && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                   ? __ldg(x + (int64_t)b0 * HW + (int64_t)gy * p.W + gx) : 0.f;
      }
    };
    float nxt[2];
    load_patch(blockIdx.x, nxt);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(bar_aempty + s * 8, ph ^ 1u);            // the UMMAs that read this A buffer are done
      float* patch = reinterpret_cast<float*>(smem_raw + L::PATCH_OFF) + s * L::PATCH_FLOATS;
      patch[m] = nxt[0];
      if (m + 128 < PH * PW) patch[m + 128] = nxt[1];
      load_patch(tile + gridDim.x, nxt);                 // in flight during the build below
      named_bar_sync(5, 128);
      // k = 0..8: hi = the 16-bit rounding of x_tap; k = 9..17: the rounding of (x_tap - hi)
      uint16_t kv[24];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) {
        const float v = patch[(ty + t9 / 3) * PW + tx + t9 % 3];
        if (p.f16) { kv[t9] = cvt16<true>(v); kv[9 + t9] = cvt16<true>(v - cvt16_to_f32<true>(kv[t9])); }
        else { kv[t9] = cvt16<false>(v); kv[9 + t9] = cvt16<false>(v - cvt16_to_f32<false>(kv[t9])); }
      }
#pragma unroll
      for (int j = 18; j < 24; ++j) kv[j] = 0;
      const uint32_t arow = smem_base + s * L::A_BYTES + m * 128;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pk[j] = (uint32_t)kv[c * 8 + 2 * j] | ((uint32_t)kv[c * 8 + 2 * j + 1] << 16);
        sts128_u32(arow + (((c ^ (m & 7)) & 7) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_afull + s * 8);
      named_bar_sync(5, 128);                              // patch[s ^ ...] reuse: everyone done reading before the next fill
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry point (no libcuda link)
// ------------------------------------------------------------------------------------
// TMA descriptors: encoded through the launch context (ctx.cuh) — cached when one is bound to the calling thread.
static int tiled_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                     CUtensorMapL2promotion promo, const char* what) {
  TensorMapSpec s{};
  s.ptr = ptr; s.rank = rank; s.dtype = (int)CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; s.swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  s.l2promo = (int)promo; s.oob = (int)CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE;
  for (int i = 0; i < rank; ++i) { s.dims[i] = dims[i]; s.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s.strides[i] = strides[i];
  return tensor_map(m, s, what);
}

// NHWC bf16 activation [B][H][W][C]: box = 64 channels x TW x TH x TB, 128 B swizzle, zero OOB fill
static int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int TW, int TH, int TB) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[4] = {(uint32_t)TC_BK, (uint32_t)TW, (uint32_t)TH, (uint32_t)TB};
  return tiled_map(m, ptr, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "activation");
}
// packed weights [N][K] bf16, K fastest: box = 64 x BN
static int make_w_map(CUtensorMap* m, const void* ptr, int N, int K, int BN) {
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t strides[1] = {(uint64_t)K * 2};
  const uint32_t box[2] = {(uint32_t)TC_BK, (uint32_t)BN};
  return tiled_map(m, ptr, 2, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weights");
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <int BN, int STAGES, int MINB, int NSTG = 1, int RESW = 0, bool TPAIR = false, bool STATS = false>
static int launch_conv_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& wm, const CUtensorMap* ym,
                          const ConvTcParams& p, const float* bias, void* y, void* y_pool, int64_t grid,
                          cudaStream_t st) {
  using L = ConvTcSmem<BN, STAGES, NSTG, RESW>;
  static_assert(MINB * (L::DYN_BYTES + 1024) <= 228 * 1024 && L::DYN_BYTES <= 227 * 1024, "shared memory budget");
  auto kern = conv_tc_kernel<BN, STAGES, MINB, NSTG, RESW, TPAIR, STATS>;
  { const int rc_ = set_max_dyn_smem(reinterpret_cast<const void*>(kern), L::DYN_BYTES); if (rc_) return rc_; }
  grid = std::min<int64_t>(grid, (int64_t)sm_count() * MINB);   // persistent: MINB CTAs per SM
  kern<<<(unsigned)grid, TC_THREADS, L::DYN_BYTES, st>>>(a0, a1, wm, ym[0], ym[1], ym[2], ym[3], p, bias,
                                                         reinterpret_cast<__nv_bfloat16*>(y),
                                                         reinterpret_cast<__nv_bfloat16*>(y_pool));
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

// output of ConvTranspose2d k2 s2, phase (i,j): the pixels (2h+i, 2w+j) of y[B][2H][2W][Cout] seen as a
// strided [B][H][W][Cout] tensor, so the same {64 ch, TW, TH, TB} brick store applies
// (Ho, Wo) = extents of the output TENSOR: 2H x 2W, or the skip tensor's size when Up.forward pads the upsampled map
// (F.pad, unet_parts.py:58-62: the pad of an odd extent goes to the high side, so the data sits at the origin)
static int make_convt_out_map(CUtensorMap* m, void* y, int B, int H, int W, int Cout, int ij, int TW, int TH, int TB,
                              int Ho, int Wo) {
  const int i = ij >> 1, j = ij & 1;
  char* base = reinterpret_cast<char*>(y) + ((int64_t)i * Wo + j) * Cout * 2;
  const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)2 * Cout * 2, (uint64_t)2 * Wo * Cout * 2, (uint64_t)Ho * Wo * Cout * 2};
  const uint32_t box[4] = {(uint32_t)TC_BK, (uint32_t)TW, (uint32_t)TH, (uint32_t)TB};
  return tiled_map(m, base, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "convT output phase");
}

// rows of parity i of the ConvTranspose2d output y[B][2H][2W][Cout] as a [B][H][2W][Cout] tensor (row stride = two
// output rows): one box {64 ch, 2 TW, TH, TB} covers both column parities of a tile (convt_pair_epilogue)
static int make_convt_pair_map(CUtensorMap* m, void* y, int B, int H, int W, int Cout, int i, int TW, int TH, int TB,
                               int Ho, int Wo) {
  char* base = reinterpret_cast<char*>(y) + (int64_t)i * Wo * Cout * 2;
  const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)(2 * W), (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)2 * Wo * Cout * 2, (uint64_t)Ho * Wo * Cout * 2};
  const uint32_t box[4] = {(uint32_t)TC_BK, (uint32_t)(2 * TW), (uint32_t)TH, (uint32_t)TB};
  return tiled_map(m, base, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "convT output rows of one parity");
}

// Tile configuration (measured per layer on B200, scripts/time_convs.py, profiles/r01_conv_variants.txt,
// profiles/r02_experiments.txt):
//   Cout % 256 == 0, convT : BN = 256, one persistent CTA per SM, 4-stage ring (A traffic per flop halves;
//                            up to ~1.5 PFLOP/s on the 256..1024-channel layers)
//   3x3, Cout 64 / 128 on images >= 16 rows : the row-shift kernel (conv_rs_kernel), weights resident at Cin = 64
//   otherwise              : BN = 128 / 64 with TWO persistent CTAs per SM (3 / 4 stages each): two
//                            independent MMA issuers hide each other's barrier round trips.

// Cin = 1, Cout = 64, W >= 16, H >= 8: tensor-core first layer; returns PMU_ERR_UNSUPPORTED otherwise (the caller
// falls back to the CUDA-core stencil of layers_bf16.cu)
int conv_first_tc_launch(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int relu,
                         int f16, cudaStream_t st) {
  int cc_major = 0;
  { const int rc_ = device_cc_major(&cc_major); if (rc_) return rc_; }
  if (cc_major != 10 || W < 16 || H < 8) return PMU_ERR_UNSUPPORTED;
  ConvTcParams p;
  p.B = B; p.H = H; p.W = W; p.C0 = 64; p.C1 = 0; p.Cout = 64; p.ntaps = 9; p.relu = relu;
  p.pool_mode = -1; p.tma_store = 1; p.f16 = f16 ? 1 : 0;
  p.TW = 16; p.TH = 8; p.TB = 1;
  p.tiles_w = cdiv(W, 16); p.tiles_h = cdiv(H, 8); p.tiles_b = B; p.n_tiles = 1;
  const int64_t tiles = (int64_t)p.tiles_w * p.tiles_h * B;
  if (tiles >= (1ll << 31)) return PMU_ERR_UNSUPPORTED;
  CUtensorMap ym;
  int rc = make_act_map(&ym, y, B, H, W, 64, 16, 8, 1);
  if (rc) return rc;
  { const int rc_ = set_max_dyn_smem(reinterpret_cast<const void*>(conv_first_tc_kernel), ConvFirstSmem::DYN_BYTES); if (rc_) return rc_; }
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sm_count() * 2);
  conv_first_tc_kernel<<<grid, CF_THREADS, ConvFirstSmem::DYN_BYTES, st>>>(ym, p, x, w, bias);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

}  // namespace pmu

using namespace pmu;

static int conv_gemm_impl(const void* x0, int C0, const void* x1, int C1, const void* wpack,
                          const float* bias, void* y, void* y_pool, int pool_mode, int B, int H, int W, int Cout,
                          int ntaps, int relu, int f16, int out_h, int out_w, void* stream, int bias_bstride = 0,
                          double* stats = nullptr) {
  PMU_CHECK_ARG(x0 && wpack && (y || y_pool), "pmu_conv_gemm_bf16: null pointer");
  const int Ho = out_h > 0 ? out_h : 2 * H, Wo = out_w > 0 ? out_w : 2 * W;     // convT output tensor extents
  PMU_CHECK_ARG(ntaps == 4 ? (Ho >= 2 * H && Wo >= 2 * W) : (out_h <= 0 && out_w <= 0),
                "pmu_conv_gemm_bf16: out_h / out_w are the (padded) output extents of the transposed convolution, >= 2H x 2W");
  PMU_CHECK_ARG(ntaps == 9 || ntaps == 4 || ntaps == 1, "pmu_conv_gemm_bf16: ntaps must be 9, 4 or 1 (got %d)", ntaps);
  PMU_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cout > 0 && C0 > 0 && C1 >= 0, "pmu_conv_gemm_bf16: bad shape");
  PMU_CHECK_ARG(C1 == 0 || x1, "pmu_conv_gemm_bf16: C1 > 0 needs x1");
  PMU_CHECK_ARG(!(ntaps == 4 && (C1 != 0 || relu)), "pmu_conv_gemm_bf16: convT takes one source and no ReLU");
  PMU_CHECK_SUPPORTED(C0 % 64 == 0 && C1 % 64 == 0 && Cout % 64 == 0,
                      "pmu_conv_gemm_bf16: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  PMU_CHECK_ARG(aligned16(x0) && aligned16(wpack) && (!y || aligned16(y)) && (!x1 || aligned16(x1)),
                "pmu_conv_gemm_bf16: pointers must be 16-byte aligned");
  int cc_major = 0;
  { const int rc_ = device_cc_major(&cc_major); if (rc_) return rc_; }
  PMU_CHECK_SUPPORTED(cc_major == 10, "pmu_conv_gemm_bf16: needs an sm_100 device (tcgen05/TMEM); found cc %d.x", cc_major);

  ConvTcParams p;
  p.B = B; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.Cout = Cout; p.ntaps = ntaps; p.relu = relu;
  p.pool_mode = y_pool ? pool_mode : -1;
  p.f16 = f16 ? 1 : 0;
  p.TW = std::min(16, pow2ceil(W));
  p.TH = std::min(TC_BM / p.TW, pow2ceil(H));
  p.TB = TC_BM / (p.TW * p.TH);
  p.tiles_w = cdiv(W, p.TW); p.tiles_h = cdiv(H, p.TH); p.tiles_b = cdiv(B, p.TB);
  const int Cin = C0 + C1;
  const int Ntot = (ntaps == 4) ? 4 * Cout : Cout;
  const int Ktot = (ntaps == 9) ? 9 * Cin : Cin;
  int BN = (Cout % 128 == 0) ? 128 : 64;
  if (Cout % 256 == 0 || ntaps == 4) BN = 256;   // convT: Ntot = 4*Cout
  // few-tile launches (the deep layers of a batch-8 training step: 16 pixel tiles x 4 N tiles on 148 SMs, every CTA a
  // serial chain of 576 UMMAs): halve the tile so that twice as many SMs share the same chain
  const bool halved = BN == 256 && ntaps != 4 &&
                      (int64_t)p.tiles_w * p.tiles_h * p.tiles_b * (Ntot / 256) * 2 <= (int64_t)sm_count();
  if (halved) BN = 128;
  p.n_tiles = Ntot / BN;
  if (y_pool) {
    PMU_CHECK_ARG(pool_mode == PMU_POOL_MAX || pool_mode == PMU_POOL_AVG_CEIL, "pmu_conv_gemm_pool_bf16: unknown pool mode %d", pool_mode);
    PMU_CHECK_SUPPORTED(ntaps != 4 && H % 2 == 0 && W % 2 == 0,
                        "pmu_conv_gemm_pool_bf16: fused pooling needs even H, W (got %dx%d)", H, W);
    PMU_CHECK_ARG(aligned16(y_pool), "pmu_conv_gemm_pool_bf16: y_pool must be 16-byte aligned");
  }
  // small-N 3x3 layers on images >= 16 rows: row-shift kernel (see conv_rs_kernel)
  const bool rs = ntaps == 9 && BN <= 128 && !halved && H >= 16 && W >= 8 && (!y_pool || (H % 2 == 0 && W % 2 == 0));
  if (rs) {
    p.TW = 8; p.TH = 16; p.TB = 1;
    p.tiles_w = cdiv(W, 8); p.tiles_h = cdiv(H, 16); p.tiles_b = B;
  } else if (y_pool) {
    PMU_CHECK_SUPPORTED(p.TW == 16 && p.TH == 8, "pmu_conv_gemm_pool_bf16: fused pooling needs W >= 16, H >= 8 (got %dx%d)", H, W);
  }
  const int64_t grid = (int64_t)p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles;
  PMU_CHECK_ARG(grid > 0 && grid < (1ll << 31), "pmu_conv_gemm_bf16: grid too large");
  p.bias_bstride = bias_bstride;
  p.stats = stats;
  PMU_CHECK_ARG(!stats || (y != nullptr && ntaps != 4), "pmu_conv_gemm_bnstats_bf16: statistics need a stored output of a 3x3 / 1x1 convolution");
  PMU_CHECK_SUPPORTED(bias_bstride == 0 || (p.TB == 1 && ntaps == 1 && bias),
                      "pmu_conv1x1_slicebias_bf16: a per-image bias needs images of at least 128 pixels (got %dx%d)", H, W);

  CUtensorMap a0, a1, wm;
  const int a_th = rs ? p.TH + 2 : p.TH;      // row-shift kernel: the A box carries the two halo rows
  int rc = make_act_map(&a0, x0, B, H, W, C0, p.TW, a_th, p.TB);
  if (rc) return rc;
  if (C1 > 0) { rc = make_act_map(&a1, x1, B, H, W, C1, p.TW, a_th, p.TB); if (rc) return rc; }
  else a1 = a0;
  rc = make_w_map(&wm, wpack, Ntot, Ktot, BN);
  if (rc) return rc;
  // output tensor maps for the TMA-store epilogue
  CUtensorMap ym[4];
  p.tma_store = (y != nullptr) ? 1 : 0;
  const bool tpair = BN == 256 && ntaps == 4 && Cout == 64 && Cin <= 128 && y != nullptr;   // paired-phase stores, resident weights
  if (p.tma_store) {
    if (tpair) {
      for (int i = 0; i < 2; ++i) {
        rc = make_convt_pair_map(&ym[i], y, B, H, W, Cout, i, p.TW, p.TH, p.TB, Ho, Wo);
        if (rc) return rc;
      }
      ym[2] = ym[3] = ym[0];
    } else if (ntaps == 4) {
      for (int ij = 0; ij < 4; ++ij) {
        rc = make_convt_out_map(&ym[ij], y, B, H, W, Cout, ij, p.TW, p.TH, p.TB, Ho, Wo);
        if (rc) return rc;
      }
    } else {
      rc = make_act_map(&ym[0], y, B, H, W, Cout, p.TW, p.TH, p.TB);
      if (rc) return rc;
      ym[1] = ym[2] = ym[3] = ym[0];
    }
  } else {
    ym[0] = ym[1] = ym[2] = ym[3] = a0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (rs) {
    auto launch_rs = [&](auto kern, int dyn) -> int {
      { const int rc_ = set_max_dyn_smem(reinterpret_cast<const void*>(kern), dyn); if (rc_) return rc_; }
      const unsigned g = (unsigned)std::min<int64_t>(grid, sm_count());
      kern<<<g, TC_THREADS, dyn, st>>>(a0, a1, wm, ym[0], p, bias, reinterpret_cast<__nv_bfloat16*>(y),
                                       reinterpret_cast<__nv_bfloat16*>(y_pool));
      PMU_LAUNCH_CHECK();
      return PMU_OK;
    };
    if (stats) {     // training: BatchNorm statistics in the epilogue
      if (BN == 64 && Cin == 64) return launch_rs(conv_rs_kernel<64, 6, 1, true>, ConvRsSmem<64, 6, 1>::DYN_BYTES);
      if (BN == 128 && Cin == 64) return launch_rs(conv_rs_kernel<128, 3, 1, true>, ConvRsSmem<128, 3, 1>::DYN_BYTES);
      if (BN == 64) return launch_rs(conv_rs_kernel<64, 4, 0, true>, ConvRsSmem<64, 4, 0>::DYN_BYTES);
      return launch_rs(conv_rs_kernel<128, 3, 0, true>, ConvRsSmem<128, 3, 0>::DYN_BYTES);
    }
    if (BN == 64 && Cin == 64) return launch_rs(conv_rs_kernel<64, 6, 1>, ConvRsSmem<64, 6, 1>::DYN_BYTES);      // 64 -> 64: weights resident
    if (BN == 128 && Cin == 64) return launch_rs(conv_rs_kernel<128, 3, 1>, ConvRsSmem<128, 3, 1>::DYN_BYTES);   // 64 -> 128: weights resident
    if (BN == 64) return launch_rs(conv_rs_kernel<64, 4, 0>, ConvRsSmem<64, 4, 0>::DYN_BYTES);
    return launch_rs(conv_rs_kernel<128, 3, 0>, ConvRsSmem<128, 3, 0>::DYN_BYTES);
  }
  if (stats) {
    if (BN == 256) return launch_conv_tc<256, 4, 1, 2, 0, false, true>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
    if (BN == 128) return launch_conv_tc<128, 3, 2, 1, 0, false, true>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
    return launch_conv_tc<64, 4, 2, 1, 0, false, true>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
  }
  if (tpair) return launch_conv_tc<256, 5, 1, 4, 2, true>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
  if (BN == 256) return launch_conv_tc<256, 4, 1, 2>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
  if (BN == 128) return launch_conv_tc<128, 3, 2>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
  return launch_conv_tc<64, 4, 2>(a0, a1, wm, ym, p, bias, y, y_pool, grid, st);
}

extern "C" int pmu_conv_gemm_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack,
                                  const float* bias, void* y, int B, int H, int W, int Cout, int ntaps,
                                  int relu, int f16, int out_h, int out_w, void* stream) {
  return conv_gemm_impl(x0, C0, x1, C1, wpack, bias, y, nullptr, -1, B, H, W, Cout, ntaps, relu, f16, out_h, out_w, stream);
}

extern "C" int pmu_conv_gemm_pool_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack,
                                       const float* bias, void* y, void* y_pool, int pool_mode, int B, int H,
                                       int W, int Cout, int relu, int f16, void* stream) {
  PMU_CHECK_ARG(y_pool != nullptr, "pmu_conv_gemm_pool_bf16: y_pool is null");
  return conv_gemm_impl(x0, C0, x1, C1, wpack, bias, y, y_pool, pool_mode, B, H, W, Cout, 9, relu, f16, 0, 0, stream);
}

extern "C" int pmu_conv1x1_slicebias_bf16(const void* x, int Cin, const void* wpack, const float* bias, void* y, int B, int H,
                                          int W, int Cout, int relu, int f16, void* stream) {
  PMU_CHECK_ARG(bias != nullptr, "pmu_conv1x1_slicebias_bf16: bias is null");
  return conv_gemm_impl(x, Cin, nullptr, 0, wpack, bias, y, nullptr, -1, B, H, W, Cout, 1, relu, f16, 0, 0, stream, Cout);
}

extern "C" int pmu_conv_gemm_bnstats_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                                          void* y, double* stats, int B, int H, int W, int Cout, int ntaps, int f16,
                                          void* stream) {
  PMU_CHECK_ARG(stats != nullptr, "pmu_conv_gemm_bnstats_bf16: stats is null");
  return conv_gemm_impl(x0, C0, x1, C1, wpack, bias, y, nullptr, -1, B, H, W, Cout, ntaps, 0, f16, 0, 0, stream, 0, stats);
}
