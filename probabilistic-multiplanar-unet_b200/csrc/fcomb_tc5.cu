// K3 + K4 fused on the 5th-gen tensor cores — Fcomb over N latent samples (tcgen05 + TMEM),
// softmax, per-pixel sum / sum-of-squares.  Per-sample logits never reach HBM.
//
// Replaces Fcomb.forward (probabilistic_unet.py:155-181) called once per sample from
// ProbabilisticUnet.sample (:225-240), the softmax of eval.py:157 and the sample loop of
// eval.py:146-154.
//
// Why tcgen05: the register-chained mma.sync version (fcomb_tc.cu) is pinned at ~280 TFLOP/s —
// the legacy warp-level MMA path of sm_100 — no matter how it is scheduled.  Here every layer is
// a UMMA, and the kernel is organised around the real bottleneck of a 64-wide MLP chain, the
// MMA -> epilogue -> MMA round-trip LATENCY (ncu: tensor pipe 33 %, issue 46 %, everything waits):
//   * tile = 128 pixels (UMMA M).  One CTA runs four warpgroups (128 threads, thread = pixel row),
//     each on its own tile and each with TWO samples in flight (ping-pong), i.e. 8 independent
//     layer chains per SM; warpgroup w has its own issuer thread (warp 16 + w).
//   * layer 0 is split (App. A): h0 = relu(W0f f + (W0z z_n + b0)).  The feature tile F (TMA,
//     128-byte swizzle) stays in smem for all N samples; the per-sample vector zb_n = W0z z_n + b0
//     — like the constant biases of the other layers — is PRE-FILLED into the TMEM accumulator by
//     the epilogue threads (tcgen05.st) and every UMMA accumulates on top of it: exact fp32 bias,
//     no bias instructions on the critical path after the MMA, no extra operand tiles.
//   * activations go TMEM -(tcgen05.ld)-> registers -(cvt.rn.relu.bf16x2: ReLU + pack in one
//     instruction per two elements)-> smem in the K-major 128B-swizzled layout -> next UMMA's A.
//   * head 64 -> C is a UMMA with N = 16 into the same accumulator columns; softmax is
//     per-thread (a thread owns a pixel's classes); sum p / sum p^2 accumulate in registers.
#include <cudaTypedefs.h>

#include "pmu_common.cuh"
#include "sm100_ptx.cuh"

namespace pmu {

using namespace ptx;

constexpr int F5_F = 64;            // feature width
constexpr int F5_WG = 4;            // warpgroups (tiles in flight) per CTA
constexpr int F5_SLOTS = 2;         // samples in flight per warpgroup
constexpr int F5_THREADS = F5_WG * 128 + F5_WG * 32;   // 4 warpgroups + one issuer warp per warpgroup
constexpr int F5_NS = 16;           // samples per bias group held in smem
constexpr int F5_MAXL = 16;
constexpr int F5_MAXC = 8;

// shared memory map (operand tiles 1024 B aligned, rows of 128 B = 64 bf16, 128B swizzle)
constexpr int F5_TILE = 128 * 128;                 // 16 KB: [128 rows][64 k]
constexpr int F5_WT = 64 * 128;                    // 8 KB:  [64 rows][64 k]
constexpr int F5_OFF_W0 = 0;                       // W0f
constexpr int F5_OFF_WM = F5_OFF_W0 + F5_WT;       // up to 2 mid layers
constexpr int F5_OFF_WL = F5_OFF_WM + 2 * F5_WT;   // head [16 rows][64 k] (2 KB)
constexpr int F5_OFF_WG = F5_OFF_WL + 2048;        // per warpgroup: F tile, H tile x SLOTS
constexpr int F5_WG_BYTES = (1 + F5_SLOTS) * F5_TILE;
constexpr int F5_OFF_ZB = F5_OFF_WG + F5_WG * F5_WG_BYTES;   // fp32 zb[F5_NS][64]
constexpr int F5_OFF_BM = F5_OFF_ZB + F5_NS * F5_F * 4;      // fp32 bmid[2][64]
constexpr int F5_OFF_BL = F5_OFF_BM + 2 * F5_F * 4;          // fp32 blast[16]
constexpr int F5_OFF_BAR = F5_OFF_BL + 64;
constexpr int F5_NBAR = F5_WG * (2 * F5_SLOTS + 1);          // ready[slot], acc[slot], tma  per warpgroup
constexpr int F5_OFF_TPTR = F5_OFF_BAR + F5_NBAR * 8;
constexpr int F5_SMEM = F5_OFF_TPTR + 16;
static_assert(F5_OFF_WG % 1024 == 0, "operand tiles must be 1024 B aligned");
static_assert(F5_SMEM <= 227 * 1024, "shared memory budget");

struct Fcomb5Params {
  int N, L, C, nmid;
  int64_t HW;
  int quads_per_cta;     // tile quads each CTA walks through
};

// byte offset of element (row, k) inside a K-major 128B-swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ void st_bf16(uint8_t* tile, int row, int k, float v) {
  *reinterpret_cast<__nv_bfloat16*>(tile + sw128_off(row, k)) = __float2bfloat16(v);
}
// ReLU + round-to-nearest bf16 pack of two fp32 values in ONE instruction (lo -> bits 0..15)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> 32 lanes x 16 consecutive fp32 TMEM columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// pre-fill NCOLS accumulator columns of this thread's lane with a bias vector read from smem
template <int NCOLS>
__device__ __forceinline__ void prefill_bias(uint32_t taddr, uint32_t bias_smem) {
#pragma unroll 1
  for (int c0 = 0; c0 < NCOLS; c0 += 16) {      // not unrolled: keeps only 16 bias registers live
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = lds128f(bias_smem + (c0 + 4 * j) * 4);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    tmem_st_32x16(taddr + c0, v);
  }
  tmem_st_wait();
}

// one dense layer on top of the pre-filled bias: D[tmem] += A[128 x 64] * W[NOUT x 64]^T  (4 UMMAs)
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile, uint32_t idesc) {
  const uint64_t ad = umma_smem_desc_sw128(a_tile), wd = umma_smem_desc_sw128(w_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, 1u);
}

template <int CMAX>
__global__ void __launch_bounds__(F5_THREADS, 1)
fcomb_tc5_kernel(const __grid_constant__ CUtensorMap tmF, const Fcomb5Params p, const float* __restrict__ mu,
                 const float* __restrict__ sigma, const float* __restrict__ eps, const float* __restrict__ w0,
                 const float* __restrict__ b0, const float* __restrict__ wmid, const float* __restrict__ bmid,
                 const float* __restrict__ wlast, const float* __restrict__ blast,
                 float* __restrict__ slice_sums) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  uint8_t* sgen = smem_raw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int N = p.N, L = p.L, C = p.C, nmid = p.nmid;
  const int64_t HW = p.HW;
  if ((sbase & 1023u) != 0) __trap();   // UMMA/TMA tiles need 1024 B alignment

  // barriers: kind 0 = ready[slot] (128 arrivals), 1 = acc[slot] (tcgen05.commit), 2 = tma
  auto bar_ready = [&](int w, int s) { return sbase + F5_OFF_BAR + ((w * (2 * F5_SLOTS + 1)) + s) * 8; };
  auto bar_acc = [&](int w, int s) { return sbase + F5_OFF_BAR + ((w * (2 * F5_SLOTS + 1)) + F5_SLOTS + s) * 8; };
  auto bar_tma = [&](int w) { return sbase + F5_OFF_BAR + ((w * (2 * F5_SLOTS + 1)) + 2 * F5_SLOTS) * 8; };
  volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(sgen + F5_OFF_TPTR);

  // ---------------- one-time setup: barriers, TMEM, weight tiles, constant biases ----------------
  if (tid == 0) {
    prefetch_tensormap(&tmF);
    for (int w = 0; w < F5_WG; ++w) {
      for (int s = 0; s < F5_SLOTS; ++s) { mbar_init(bar_ready(w, s), 128); mbar_init(bar_acc(w, s), 1); }
      mbar_init(bar_tma(w), 1);
    }
    fence_barrier_init();
  }
  if (warp == F5_WG * 4) tmem_alloc<512>(sbase + F5_OFF_TPTR);   // first issuer warp owns the allocation
  for (int i = tid; i < 2048 / 16; i += F5_THREADS) reinterpret_cast<uint4*>(sgen + F5_OFF_WL)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < F5_F * F5_F; i += F5_THREADS) {
    const int o = i >> 6, k = i & 63;
    st_bf16(sgen + F5_OFF_W0, o, k, __ldg(w0 + (int64_t)o * (F5_F + L) + k));
    for (int m = 0; m < nmid; ++m) st_bf16(sgen + F5_OFF_WM + m * F5_WT, o, k, __ldg(wmid + (int64_t)m * F5_F * F5_F + i));
  }
  for (int i = tid; i < C * F5_F; i += F5_THREADS) st_bf16(sgen + F5_OFF_WL, i >> 6, i & 63, __ldg(wlast + i));
  float* bm_s = reinterpret_cast<float*>(sgen + F5_OFF_BM);
  float* bl_s = reinterpret_cast<float*>(sgen + F5_OFF_BL);
  float* zb_s = reinterpret_cast<float*>(sgen + F5_OFF_ZB);
  for (int i = tid; i < nmid * F5_F; i += F5_THREADS) bm_s[i] = __ldg(bmid + i);
  for (int i = tid; i < 16; i += F5_THREADS) bl_s[i] = (i < C) ? __ldg(blast + i) : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tptr;

  const int64_t tiles = (HW + 127) / 128;
  const int64_t quads = (tiles + F5_WG - 1) / F5_WG;
  const int64_t q_lo = (int64_t)blockIdx.x * p.quads_per_cta;
  const int64_t q_hi = (q_lo + p.quads_per_cta < quads) ? q_lo + p.quads_per_cta : quads;

  // barrier phase parities: bit s of phr / pha = slot s
  uint32_t phr = 0, pha = 0, pht = 0;

  for (int n0 = 0; n0 < N; n0 += F5_NS) {
    const int ng = (N - n0 < F5_NS) ? N - n0 : F5_NS;
    // ---- per-sample layer-0 bias vectors zb_n = W0z z_n + b0 of this sample group (fp32, smem) ----
    __syncthreads();                                  // everyone is done with the previous group's vectors
    for (int i = tid; i < ng * F5_F; i += F5_THREADS) {
      const int n = i >> 6, o = i & 63;
      float s = __ldg(b0 + o);
      for (int l = 0; l < L; ++l) {
        // z = mu + sigma * eps   (Normal.rsample, probabilistic_unet.py:233)
        const float z = __fadd_rn(__ldg(mu + (int64_t)b * L + l),
                                  __fmul_rn(__ldg(sigma + (int64_t)b * L + l), __ldg(eps + ((int64_t)b * N + n0 + n) * L + l)));
        s = fmaf(__ldg(w0 + (int64_t)o * (F5_F + L) + F5_F + l), z, s);
      }
      zb_s[i] = s;
    }
    fence_proxy_async_smem();                         // weight tiles (generic proxy) -> visible to the tensor core
    __syncthreads();
    // samples of this group handled by slot s: n = s, s + SLOTS, ...  (identical on both sides)
    const int rounds = (ng + F5_SLOTS - 1) / F5_SLOTS;

    if (warp >= F5_WG * 4) {
      // ============ issuer of warpgroup w (one elected thread) ============
      if (elect_one()) {
        const int w = warp - F5_WG * 4;
        constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64), idesc16 = umma_idesc_bf16(128, 16);
        const uint32_t sW0 = sbase + F5_OFF_W0, sWM = sbase + F5_OFF_WM, sWL = sbase + F5_OFF_WL;
        const uint32_t sF = sbase + F5_OFF_WG + w * F5_WG_BYTES;
        for (int64_t q = q_lo; q < q_hi; ++q) {
          const int64_t t = q * F5_WG + w;
          if (t >= tiles) continue;
          bool have_f = false;
          for (int r = 0; r < rounds; ++r) {
            for (int layer = 0; layer <= nmid + 1; ++layer) {        // 0: layer 0 (A = F); 1..nmid: mid; nmid+1: head
              for (int s = 0; s < F5_SLOTS; ++s) {
                if (r * F5_SLOTS + s >= ng) continue;
                const uint32_t t_acc = tmem_base + (w * F5_SLOTS + s) * 64;
                const uint32_t sH = sF + (1 + s) * F5_TILE;
                mbar_wait(bar_ready(w, s), (phr >> s) & 1u); phr ^= 1u << s;   // bias pre-filled / H written
                tcgen05_fence_after();
                if (!have_f) {
                  // first use of this tile: the warpgroup's first arrival for it also says that the previous
                  // tile is completely finished (program order), so F may be overwritten now
                  mbar_arrive_expect_tx(bar_tma(w), F5_TILE);
                  tma_load_2d(sF, &tmF, bar_tma(w), 0, (int)((int64_t)b * HW + t * 128));
                  mbar_wait(bar_tma(w), pht); pht ^= 1u;
                  tcgen05_fence_after();
                  have_f = true;
                }
                if (layer == 0) issue_layer(t_acc, sF, sW0, idesc64);
                else if (layer <= nmid) issue_layer(t_acc, sH, sWM + (layer - 1) * F5_WT, idesc64);
                else issue_layer(t_acc, sH, sWL, idesc16);
                umma_commit(bar_acc(w, s));
              }
            }
          }
        }
      }
      __syncwarp();
    } else {
      // ============ warpgroup w: one tile, F5_SLOTS samples in flight ============
      const int w = warp >> 2, q4 = warp & 3;
      const int row = q4 * 32 + lane;                            // TMEM lane == pixel row of the tile
      const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
      // the 8 swizzled 16-byte slots of this thread's row inside an H tile (loop invariant)
      uint32_t hoff[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) hoff[c] = row * 128 + (((c ^ (row & 7)) & 7) << 4);
      const uint32_t sHbase = sbase + F5_OFF_WG + w * F5_WG_BYTES + F5_TILE;
      const uint32_t sZB = sbase + F5_OFF_ZB, sBM = sbase + F5_OFF_BM, sBL = sbase + F5_OFF_BL;

      for (int64_t q = q_lo; q < q_hi; ++q) {
        const int64_t t = q * F5_WG + w;
        if (t >= tiles) continue;
        const int64_t pix = t * 128 + row;
        float s1[CMAX], s2[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) s1[c] = s2[c] = 0.f;
        // prologue: pre-fill the layer-0 bias of the first sample of every slot
#pragma unroll
        for (int s = 0; s < F5_SLOTS; ++s) {
          if (s < ng) {
            prefill_bias<64>(tmem_base + (w * F5_SLOTS + s) * 64 + lane_off, sZB + s * F5_F * 4);
            tcgen05_fence_before();
            mbar_arrive(bar_ready(w, s));
          }
        }
        for (int r = 0; r < rounds; ++r) {
          for (int layer = 0; layer <= nmid; ++layer) {
#pragma unroll
            for (int s = 0; s < F5_SLOTS; ++s) {
              if (r * F5_SLOTS + s >= ng) continue;
              const uint32_t t_acc = tmem_base + (w * F5_SLOTS + s) * 64 + lane_off;
              const uint32_t sH = sHbase + s * F5_TILE;
              mbar_wait(bar_acc(w, s), (pha >> s) & 1u); pha ^= 1u << s;   // this layer's accumulator is complete
              tcgen05_fence_after();
#pragma unroll
              for (int half = 0; half < 2; ++half) {             // 2 x 32 accumulator columns (register budget: 96)
                uint32_t rr[32];
                tmem_ld_32x32(t_acc + half * 32, rr);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c)                      // chunks of 8 channels = 16 B each
                  sts128(sH + hoff[half * 4 + c],
                         pack_relu_bf16(__uint_as_float(rr[c * 8 + 0]), __uint_as_float(rr[c * 8 + 1])),
                         pack_relu_bf16(__uint_as_float(rr[c * 8 + 2]), __uint_as_float(rr[c * 8 + 3])),
                         pack_relu_bf16(__uint_as_float(rr[c * 8 + 4]), __uint_as_float(rr[c * 8 + 5])),
                         pack_relu_bf16(__uint_as_float(rr[c * 8 + 6]), __uint_as_float(rr[c * 8 + 7])));
              }
              // bias of the layer that will run next on this accumulator
              if (layer < nmid) prefill_bias<64>(t_acc, sBM + layer * F5_F * 4);
              else prefill_bias<16>(t_acc, sBL);
              fence_proxy_async_smem();                          // H (generic proxy) -> async proxy
              tcgen05_fence_before();
              mbar_arrive(bar_ready(w, s));                      // H ready, accumulator drained + pre-filled
            }
          }
          // ---- head logits -> softmax -> accumulate; then hand the slot its next sample ----
#pragma unroll
          for (int s = 0; s < F5_SLOTS; ++s) {
            const int n = r * F5_SLOTS + s;
            if (n >= ng) continue;
            const uint32_t t_acc = tmem_base + (w * F5_SLOTS + s) * 64 + lane_off;
            mbar_wait(bar_acc(w, s), (pha >> s) & 1u); pha ^= 1u << s;
            tcgen05_fence_after();
            uint32_t hr[16];
            tmem_ld_32x16(t_acc, hr);
            tmem_ld_wait();
            if (n + F5_SLOTS < ng) {                             // next sample of this slot: layer-0 bias
              prefill_bias<64>(t_acc, sZB + (n + F5_SLOTS) * F5_F * 4);
              tcgen05_fence_before();
              mbar_arrive(bar_ready(w, s));
            }
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c < C) mx = fmaxf(mx, __uint_as_float(hr[c]));
            float e[CMAX], den = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) { e[c] = (c < C) ? __expf(__uint_as_float(hr[c]) - mx) : 0.f; den += e[c]; }
            const float inv = __fdividef(1.f, den);
#pragma unroll
            for (int c = 0; c < CMAX; ++c) { const float pr = e[c] * inv; s1[c] += pr; s2[c] = fmaf(pr, pr, s2[c]); }
          }
        }
        if (pix < HW) {
          float* o1 = slice_sums + ((int64_t)b * 2 + 0) * C * HW + pix;
          float* o2 = slice_sums + ((int64_t)b * 2 + 1) * C * HW + pix;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              if (n0 == 0) { o1[(int64_t)c * HW] = s1[c]; o2[(int64_t)c * HW] = s2[c]; }
              else { o1[(int64_t)c * HW] += s1[c]; o2[(int64_t)c * HW] += s2[c]; }
            }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == F5_WG * 4) tmem_dealloc<512>(tmem_base);
}

static PFN_cuTensorMapEncodeTiled_v12000 f5_encode_fn() {
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

// implemented in fcomb_tc.cu (register-chained mma.sync version, kept as the nmid > 2 path)
extern "C" int pmu_fcomb_softmax_accum_bf16_mma(const void* feat, const float* mu, const float* sigma,
                                                const float* eps, const float* w0, const float* b0,
                                                const float* wmid, const float* bmid, const float* wlast,
                                                const float* blast, float* slice_sums, int B, int N, int L,
                                                int C, int nl, int64_t HW, void* stream);

extern "C" int pmu_fcomb_softmax_accum_bf16(const void* feat, const float* mu, const float* sigma,
                                            const float* eps, const float* w0, const float* b0,
                                            const float* wmid, const float* bmid, const float* wlast,
                                            const float* blast, float* slice_sums, int B, int N, int L,
                                            int C, int nl, int64_t HW, void* stream) {
  PMU_CHECK_ARG(feat && mu && sigma && eps && w0 && b0 && wlast && blast && slice_sums,
                "pmu_fcomb_softmax_accum_bf16: null pointer");
  PMU_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && HW > 0, "pmu_fcomb_softmax_accum_bf16: bad shape");
  PMU_CHECK_ARG(nl >= 2 && (nl == 2 || (wmid && bmid)), "pmu_fcomb_softmax_accum_bf16: no_convs_fcomb >= 2; mid weights needed for > 2");
  PMU_CHECK_SUPPORTED(L >= 1 && L <= F5_MAXL && C >= 1 && C <= F5_MAXC, "pmu_fcomb_softmax_accum_bf16: needs L <= 16, C <= 8 (got L=%d C=%d)", L, C);
  const int nmid = nl - 2;
  static int use_mma = -1;
  if (use_mma < 0) { const char* e = getenv("PMU_FCOMB_MMA_SYNC"); use_mma = (e && atoi(e)) ? 1 : 0; }
  if (nmid > 2 || use_mma || !aligned16(feat) || (int64_t)B * HW >= (1ll << 31))
    return pmu_fcomb_softmax_accum_bf16_mma(feat, mu, sigma, eps, w0, b0, wmid, bmid, wlast, blast, slice_sums,
                                            B, N, L, C, nl, HW, stream);
  int cc_major = 0, dev = 0;
  PMU_CUDA(cudaGetDevice(&dev));
  PMU_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  PMU_CHECK_SUPPORTED(cc_major == 10, "pmu_fcomb_softmax_accum_bf16: needs an sm_100 device; found cc %d.x", cc_major);

  auto fn = f5_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PMU_ERR_CUDA; }
  CUtensorMap tmF;
  cuuint64_t dims[2] = {(cuuint64_t)F5_F, (cuuint64_t)((int64_t)B * HW)};
  cuuint64_t strides[1] = {(cuuint64_t)F5_F * 2};
  cuuint32_t box[2] = {(cuuint32_t)F5_F, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tmF, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(feat), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(features) failed: %d", (int)r); return PMU_ERR_CUDA; }

  Fcomb5Params p;
  p.N = N; p.L = L; p.C = C; p.nmid = nmid; p.HW = HW;
  const int64_t tiles = (HW + 127) / 128, quads = (tiles + F5_WG - 1) / F5_WG;
  // enough CTAs for ~4 waves of the machine, but several quads per CTA to amortise the setup
  int64_t ctas_x = std::max<int64_t>(1, std::min<int64_t>(quads, (4 * (int64_t)sm_count() + B - 1) / B));
  p.quads_per_cta = (int)((quads + ctas_x - 1) / ctas_x);
  ctas_x = (quads + p.quads_per_cta - 1) / p.quads_per_cta;
  dim3 grid((unsigned)ctas_x, B);
  if (C <= 4) {
    PMU_CUDA(cudaFuncSetAttribute(fcomb_tc5_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, F5_SMEM));
    fcomb_tc5_kernel<4><<<grid, F5_THREADS, F5_SMEM, (cudaStream_t)stream>>>(tmF, p, mu, sigma, eps, w0, b0, wmid, bmid,
                                                                            wlast, blast, slice_sums);
  } else {
    PMU_CUDA(cudaFuncSetAttribute(fcomb_tc5_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, F5_SMEM));
    fcomb_tc5_kernel<8><<<grid, F5_THREADS, F5_SMEM, (cudaStream_t)stream>>>(tmF, p, mu, sigma, eps, w0, b0, wmid, bmid,
                                                                            wlast, blast, slice_sums);
  }
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
