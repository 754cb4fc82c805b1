// K3 + K4 fused on the 5th-gen tensor cores — Fcomb over N latent samples (tcgen05 + TMEM),
// softmax, per-pixel sum / sum-of-squares.  Per-sample logits never reach HBM.
//
// Replaces Fcomb.forward (probabilistic_unet.py:155-181) called once per sample from
// ProbabilisticUnet.sample (:225-240), the softmax of eval.py:157 and the sample loop of
// eval.py:146-154.
//
// Why tcgen05: the register-chained mma.sync version (fcomb_tc.cu) is pinned at ~280 TFLOP/s —
// the legacy warp-level MMA path of sm_100 — no matter how it is scheduled (three schedules,
// same 51 ms per 256^3 volume).  Here every layer is a UMMA:
//   * tile = 128 pixels (UMMA M).  One CTA runs FOUR independent tile pipelines (warpgroups
//     0..3, 128 threads each, thread = pixel row) each fed by its own issuer thread (warps 16..19), so the
//     tensor pipe always has another tile's layer to chew on while a warpgroup does its
//     TMEM -> ReLU -> bf16 -> smem epilogue.
//   * layer 0 is split (App. A): h0 = relu(W0f f + (W0z z_n + b0)).  The feature tile F (TMA,
//     128-byte swizzle) stays in smem for all N samples; the per-sample vector zb_n = W0z z_n + b0
//     enters THROUGH THE MMA as one extra K=16 block: A = E (two ones columns), B = [hi(zb_n),
//     lo(zb_n)] (bf16 hi/lo split, exact to 2^-17), so the epilogue has no bias add at all.
//     The constant biases of the 64x64 layers and of the head ride in the same way.
//   * activations go TMEM -(tcgen05.ld)-> registers -(ReLU, bf16 pack: 1 instr / element)->
//     smem in the K-major 128B-swizzled layout (chunk ^ (row & 7)) -> next UMMA's A operand.
//   * head 64 -> C is a UMMA with N = 16; softmax is per-thread (a thread owns a pixel's
//     classes); sum p and sum p^2 accumulate in registers over the N samples.
#include <cudaTypedefs.h>

#include "pmu_common.cuh"
#include "sm100_ptx.cuh"

namespace pmu {

using namespace ptx;

constexpr int F5_F = 64;            // feature width
constexpr int F5_WG = 4;            // tile pipelines (warpgroups) per CTA
constexpr int F5_THREADS = F5_WG * 128 + F5_WG * 32;   // 4 warpgroups + one issuer warp per warpgroup
constexpr int F5_NS = 16;           // samples per bias-tile group
constexpr int F5_MAXL = 16;
constexpr int F5_MAXC = 8;

// shared memory map (all tiles 1024 B aligned, rows of 128 B = 64 bf16, 128B swizzle)
constexpr int F5_TILE = 128 * 128;                 // 16 KB: [128 rows][64 k]
constexpr int F5_WT = 64 * 128;                    // 8 KB:  [64 rows][64 k]
constexpr int F5_OFF_E = 0;                        // ones tile (A operand of every bias block)
constexpr int F5_OFF_W0 = F5_OFF_E + F5_TILE;      // W0f
constexpr int F5_OFF_WM = F5_OFF_W0 + F5_WT;       // up to 2 mid layers
constexpr int F5_OFF_WL = F5_OFF_WM + 2 * F5_WT;   // head [16 rows][64 k] (2 KB used)
constexpr int F5_OFF_BT = F5_OFF_WL + 2048;        // constant-bias tile: k 0..15 mid0, 16..31 mid1, 32..47 head
constexpr int F5_OFF_ZB = F5_OFF_BT + F5_WT;       // 4 tiles: sample n -> tile n/4, k offset 16*(n%4)
constexpr int F5_OFF_WG = F5_OFF_ZB + 4 * F5_WT;   // per warpgroup: F tile, H tile
constexpr int F5_OFF_BAR = F5_OFF_WG + F5_WG * 2 * F5_TILE;
constexpr int F5_NBAR = 4 * F5_WG;                 // ready, done_acc, done_head, tma  per warpgroup
constexpr int F5_OFF_TPTR = F5_OFF_BAR + F5_NBAR * 8;
constexpr int F5_SMEM = F5_OFF_TPTR + 16 + 1024;   // + alignment slack

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
// hi/lo bf16 split of an fp32 value into k columns k0, k0+1 of row `row`
__device__ __forceinline__ void st_hilo(uint8_t* tile, int row, int k0, float v) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(tile + sw128_off(row, k0)) = hi;
  *reinterpret_cast<__nv_bfloat16*>(tile + sw128_off(row, k0 + 1)) = lo;
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
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// issue one dense layer: D[tmem] = A[128 x 64] * W[NOUT x 64]^T + E * biasblock^T   (5 UMMAs)
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile, uint32_t e_tile,
                                            uint32_t bias_tile, int bias_koff16, uint32_t idesc) {
  const uint64_t ad = umma_smem_desc_sw128(a_tile), wd = umma_smem_desc_sw128(w_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
  umma_bf16(tmem_d, umma_smem_desc_sw128(e_tile), umma_smem_desc_sw128(bias_tile) + (uint64_t)(2 * bias_koff16), idesc, 1u);
}

__global__ void __launch_bounds__(F5_THREADS, 1)
fcomb_tc5_kernel(const __grid_constant__ CUtensorMap tmF, const Fcomb5Params p, const float* __restrict__ mu,
                 const float* __restrict__ sigma, const float* __restrict__ eps, const float* __restrict__ w0,
                 const float* __restrict__ b0, const float* __restrict__ wmid, const float* __restrict__ bmid,
                 const float* __restrict__ wlast, const float* __restrict__ blast,
                 float* __restrict__ slice_sums) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int N = p.N, L = p.L, C = p.C, nmid = p.nmid;
  const int64_t HW = p.HW;
  __shared__ float zs[F5_NS * F5_MAXL];

  auto bar = [&](int kind, int w) { return sbase + F5_OFF_BAR + (kind * F5_WG + w) * 8; };  // 0 ready 1 acc 2 head 3 tma
  volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(sgen + F5_OFF_TPTR);

  // ---------------- one-time setup: barriers, TMEM, constant operand tiles ----------------
  if (tid == 0) {
    prefetch_tensormap(&tmF);
    for (int w = 0; w < F5_WG; ++w) {
      mbar_init(bar(0, w), 128);
      mbar_init(bar(1, w), 1);
      mbar_init(bar(2, w), 1);
      mbar_init(bar(3, w), 1);
    }
    fence_barrier_init();
  }
  if (warp == F5_WG * 4) tmem_alloc<512>(sbase + F5_OFF_TPTR);
  // zero E, bias tile, head tile, zb tiles (their unused k columns must read as 0)
  for (int i = tid; i < F5_TILE / 16; i += F5_THREADS) reinterpret_cast<uint4*>(sgen + F5_OFF_E)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (2048 + F5_WT + 4 * F5_WT) / 16; i += F5_THREADS)
    reinterpret_cast<uint4*>(sgen + F5_OFF_WL)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int r = tid; r < 128; r += F5_THREADS) { st_bf16(sgen + F5_OFF_E, r, 0, 1.f); st_bf16(sgen + F5_OFF_E, r, 1, 1.f); }
  for (int i = tid; i < F5_F * F5_F; i += F5_THREADS) {
    const int o = i >> 6, k = i & 63;
    st_bf16(sgen + F5_OFF_W0, o, k, __ldg(w0 + (int64_t)o * (F5_F + L) + k));
    for (int m = 0; m < nmid; ++m) st_bf16(sgen + F5_OFF_WM + m * F5_WT, o, k, __ldg(wmid + (int64_t)m * F5_F * F5_F + i));
  }
  for (int i = tid; i < C * F5_F; i += F5_THREADS) st_bf16(sgen + F5_OFF_WL, i >> 6, i & 63, __ldg(wlast + i));
  for (int i = tid; i < F5_F; i += F5_THREADS)
    for (int m = 0; m < nmid; ++m) st_hilo(sgen + F5_OFF_BT, i, 16 * m, __ldg(bmid + m * F5_F + i));
  for (int i = tid; i < C; i += F5_THREADS) st_hilo(sgen + F5_OFF_BT, i, 32, __ldg(blast + i));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tptr;

  const int64_t tiles = (HW + 127) / 128;
  const int64_t quads = (tiles + F5_WG - 1) / F5_WG;
  const int64_t q_lo = (int64_t)blockIdx.x * p.quads_per_cta;
  const int64_t q_hi = (q_lo + p.quads_per_cta < quads) ? q_lo + p.quads_per_cta : quads;

  // barrier phase parities, one bit per warpgroup (identical bookkeeping on both sides of a barrier)
  uint32_t phr = 0, pha = 0, phh = 0, pht = 0;

  for (int n0 = 0; n0 < N; n0 += F5_NS) {
    const int ng = (N - n0 < F5_NS) ? N - n0 : F5_NS;
    // ---- per-sample layer-0 bias vectors zb_n = W0z z_n + b0 for this sample group -> zb tiles ----
    __syncthreads();                                  // everyone is done with the previous group's tiles
    for (int i = tid; i < ng * L; i += F5_THREADS) {
      const int n = i / L, l = i % L;
      zs[n * F5_MAXL + l] = __fadd_rn(__ldg(mu + (int64_t)b * L + l),
                                      __fmul_rn(__ldg(sigma + (int64_t)b * L + l),
                                                __ldg(eps + ((int64_t)b * N + n0 + n) * L + l)));   // z = mu + sigma*eps
    }
    __syncthreads();
    for (int i = tid; i < ng * F5_F; i += F5_THREADS) {
      const int n = i >> 6, o = i & 63;
      float s = __ldg(b0 + o);
      for (int l = 0; l < L; ++l) s = fmaf(__ldg(w0 + (int64_t)o * (F5_F + L) + F5_F + l), zs[n * F5_MAXL + l], s);
      st_hilo(sgen + F5_OFF_ZB + (n >> 2) * F5_WT, o, 16 * (n & 3), s);
    }
    fence_proxy_async_smem();                         // generic-proxy tile writes -> visible to the tensor core
    __syncthreads();

    if (warp >= F5_WG * 4) {
      // ============ issuer of warpgroup w = warp - 16 (one elected thread; no coupling between pipelines) ============
      if (elect_one()) {
        const int w = warp - F5_WG * 4;
        constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64), idesc16 = umma_idesc_bf16(128, 16);
        const uint32_t sE = sbase + F5_OFF_E, sW0 = sbase + F5_OFF_W0, sWM = sbase + F5_OFF_WM,
                       sWL = sbase + F5_OFF_WL, sBT = sbase + F5_OFF_BT, sZB = sbase + F5_OFF_ZB;
        const uint32_t sF = sbase + F5_OFF_WG + w * 2 * F5_TILE, sH = sF + F5_TILE;
        const uint32_t t_acc = tmem_base + w * 128;
        for (int64_t q = q_lo; q < q_hi; ++q) {
          const int64_t t = q * F5_WG + w;
          if (t >= tiles) continue;
          // tile start: the warpgroup released its F / H / TMEM -> fetch the feature tile
          mbar_wait(bar(0, w), phr); phr ^= 1u;
          mbar_arrive_expect_tx(bar(3, w), F5_TILE);
          tma_load_2d(sF, &tmF, bar(3, w), 0, (int)((int64_t)b * HW + t * 128));
          mbar_wait(bar(3, w), pht); pht ^= 1u;
          tcgen05_fence_after();
          issue_layer(t_acc, sF, sW0, sE, sZB, 0, idesc64);          // layer 0, sample 0
          umma_commit(bar(1, w));
          for (int n = 0; n < ng; ++n) {
            for (int layer = 0; layer <= nmid; ++layer) {             // nmid mid layers, then the head
              mbar_wait(bar(0, w), phr); phr ^= 1u;                   // H written (and ACC drained)
              tcgen05_fence_after();
              if (layer < nmid) {
                issue_layer(t_acc, sH, sWM + layer * F5_WT, sE, sBT, layer, idesc64);
                umma_commit(bar(1, w));
              } else {
                issue_layer(t_acc + 64, sH, sWL, sE, sBT, 2, idesc16);
                umma_commit(bar(2, w));
                if (n + 1 < ng) {                                     // next sample's layer 0 needs nothing from the warpgroup
                  issue_layer(t_acc, sF, sW0, sE, sZB + ((n + 1) >> 2) * F5_WT, (n + 1) & 3, idesc64);
                  umma_commit(bar(1, w));
                }
              }
            }
          }
        }
      }
      __syncwarp();
    } else {
      // =============================== warpgroup w: one tile pipeline ===============================
      const int w = warp >> 2, q4 = warp & 3;
      const int row = q4 * 32 + lane;                            // TMEM lane == pixel row of the tile
      const uint32_t t_acc = tmem_base + w * 128 + ((uint32_t)(q4 * 32) << 16);
      // the 8 swizzled 16-byte slots of this thread's row in the H tile (loop invariant)
      uint32_t hslot[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        hslot[c] = sbase + F5_OFF_WG + w * 2 * F5_TILE + F5_TILE + row * 128 + (((c ^ (row & 7)) & 7) << 4);
      for (int64_t q = q_lo; q < q_hi; ++q) {
        const int64_t t = q * F5_WG + w;
        if (t >= tiles) continue;
        const int64_t pix = t * 128 + row;
        float s1[F5_MAXC], s2[F5_MAXC];
#pragma unroll
        for (int c = 0; c < F5_MAXC; ++c) s1[c] = s2[c] = 0.f;
        tcgen05_fence_before();
        mbar_arrive(bar(0, w));                                  // tile start: F / H / TMEM are free
        for (int n = 0; n < ng; ++n) {
          for (int layer = 0; layer <= nmid; ++layer) {
            mbar_wait(bar(1, w), pha); pha ^= 1u;          // layer's accumulator complete
            tcgen05_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_acc, r0);                            // both halves in flight, one wait
            tmem_ld_32x32(t_acc + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {                        // 8 chunks of 8 channels = 16 B each
              sts128(hslot[c],
                     pack_relu_bf16(__uint_as_float(r0[c * 8 + 0]), __uint_as_float(r0[c * 8 + 1])),
                     pack_relu_bf16(__uint_as_float(r0[c * 8 + 2]), __uint_as_float(r0[c * 8 + 3])),
                     pack_relu_bf16(__uint_as_float(r0[c * 8 + 4]), __uint_as_float(r0[c * 8 + 5])),
                     pack_relu_bf16(__uint_as_float(r0[c * 8 + 6]), __uint_as_float(r0[c * 8 + 7])));
              sts128(hslot[4 + c],
                     pack_relu_bf16(__uint_as_float(r1[c * 8 + 0]), __uint_as_float(r1[c * 8 + 1])),
                     pack_relu_bf16(__uint_as_float(r1[c * 8 + 2]), __uint_as_float(r1[c * 8 + 3])),
                     pack_relu_bf16(__uint_as_float(r1[c * 8 + 4]), __uint_as_float(r1[c * 8 + 5])),
                     pack_relu_bf16(__uint_as_float(r1[c * 8 + 6]), __uint_as_float(r1[c * 8 + 7])));
            }
            fence_proxy_async_smem();                            // H (generic proxy) -> async proxy
            tcgen05_fence_before();
            mbar_arrive(bar(0, w));                              // H ready, ACC drained
          }
          // ---- head logits -> softmax -> accumulate ----
          mbar_wait(bar(2, w), phh); phh ^= 1u;
          tcgen05_fence_after();
          uint32_t hr[16];
          tmem_ld_32x16(t_acc + 64, hr);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < F5_MAXC; ++c) if (c < C) mx = fmaxf(mx, __uint_as_float(hr[c]));
          float e[F5_MAXC], den = 0.f;
#pragma unroll
          for (int c = 0; c < F5_MAXC; ++c) { e[c] = (c < C) ? __expf(__uint_as_float(hr[c]) - mx) : 0.f; den += e[c]; }
          const float inv = __fdividef(1.f, den);
#pragma unroll
          for (int c = 0; c < F5_MAXC; ++c) { const float pr = e[c] * inv; s1[c] += pr; s2[c] = fmaf(pr, pr, s2[c]); }
        }
        if (pix < HW) {
          float* o1 = slice_sums + ((int64_t)b * 2 + 0) * C * HW + pix;
          float* o2 = slice_sums + ((int64_t)b * 2 + 1) * C * HW + pix;
#pragma unroll
          for (int c = 0; c < F5_MAXC; ++c)
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
  PMU_CUDA(cudaFuncSetAttribute(fcomb_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F5_SMEM));
  dim3 grid((unsigned)ctas_x, B);
  fcomb_tc5_kernel<<<grid, F5_THREADS, F5_SMEM, (cudaStream_t)stream>>>(tmF, p, mu, sigma, eps, w0, b0, wmid, bmid,
                                                                       wlast, blast, slice_sums);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
