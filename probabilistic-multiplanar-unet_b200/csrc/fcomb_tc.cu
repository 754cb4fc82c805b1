// K3 + K4 fused — Fcomb over N latent samples on the tensor cores, softmax, and per-pixel
// sum / sum-of-squares accumulation.  Per-sample logits never reach HBM: per pixel only
// 2*C floats leave the SM.
//
// Replaces Fcomb.forward (probabilistic_unet.py:155-181: tile z, concat, 1x1-conv MLP) called
// once per sample from ProbabilisticUnet.sample (probabilistic_unet.py:225-240), the softmax
// of eval.py:157 and the sample loop of eval.py:146-154 (App. A steps 5-6).
//
// Layer 0 is split: W0 [f; z] + b0 = W0f f + (W0z z + b0).  The feature
// GEMM u = W0f f runs ONCE per pixel tile and its fp32 accumulators stay in registers; every
// sample only adds its own 64-float bias vector zb_n = W0z z_n + b0.  The per-sample tail
// (64x64 layers, ReLU, the 64->C head) is chained register-to-register: the m16n8 fp32
// accumulator fragment of one layer, after bias + ReLU + bf16 pack, IS the m16k16 A fragment
// of the next layer, so activations never touch shared memory.  Softmax over the C classes is
// a 2-step quad shuffle.  (v1 uses warp-level mma.sync m16n8k16; the 87% of the path's FLOPs
// that sit in the 3x3 convolutions run on tcgen05 in conv_tc.cu.)
#include "pmu_common.cuh"

namespace pmu {

constexpr int FT_F = 64;          // feature width (num_filters[0] of the trainer model)
constexpr int FT_LD = 72;         // padded bf16 row stride of the weight tiles: conflict-free fragment loads
constexpr int FT_WARPS = 8;
constexpr int FT_TP = 2;          // 16-pixel m-tiles processed together by a warp (share every B fragment)
constexpr int FT_IT = 2;          // tile-pair iterations per warp  -> 8 warps * 2 * 2 * 16 = 512 pixels / block
constexpr int FT_MAXL = 16;

// D = A*B + C  (C given separately: lets the bias ride in as the initial accumulator)
__device__ __forceinline__ void mma_bf16_init(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                              float c0, float c1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mma_bf16_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// relu(x) then bf16 pack == bf16 pack then max(.,0): rounding is monotonic and keeps 0, so the
// packed HMNMX2 form is bit-identical and costs one instruction per two elements.
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  __nv_bfloat162 h2 = __hmax2(__floats2bfloat162_rn(lo, hi), __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&h2);
}

// one 64->64 layer for FT_TP 16-pixel tiles: acc[tp][nt][.] = bias + sum_k A[tp][.,k] W[nt*8+., k]
// (bias == nullptr: start from zero).  Every B fragment is loaded once and used by both tiles.
__device__ __forceinline__ void dense64(float (&acc)[FT_TP][8][4], const uint32_t (&a)[FT_TP][4][4],
                                        const __nv_bfloat16* __restrict__ Ws, const float* __restrict__ bias,
                                        int g, int t) {
  // k-step outer, n-tile inner: 16 independent accumulator chains sit between two dependent
  // MMAs, so the tensor pipe latency is covered inside one warp.
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const __nv_bfloat16* wr = Ws + (nt * 8 + g) * FT_LD + 2 * t + ks * 16;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
      if (ks == 0) {
        float2 bv = make_float2(0.f, 0.f);
        if (bias) bv = *reinterpret_cast<const float2*>(bias + nt * 8 + 2 * t);
#pragma unroll
        for (int tp = 0; tp < FT_TP; ++tp) mma_bf16_init(acc[tp][nt], a[tp][ks], b0, b1, bv.x, bv.y);
      } else {
#pragma unroll
        for (int tp = 0; tp < FT_TP; ++tp) mma_bf16_acc(acc[tp][nt], a[tp][ks], b0, b1);
      }
    }
  }
}
// ReLU + bf16 pack (bias already inside acc): accumulator fragments -> next layer's A fragments
__device__ __forceinline__ void relu_pack(uint32_t (&a)[4][4], const float (&acc)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_relu(acc[2 * ks][0], acc[2 * ks][1]);
    a[ks][1] = pack_relu(acc[2 * ks][2], acc[2 * ks][3]);
    a[ks][2] = pack_relu(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
    a[ks][3] = pack_relu(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
  }
}
// layer 0: h0 = relu(u + zb_n)
__device__ __forceinline__ void bias_relu_pack(uint32_t (&a)[4][4], const float (&u)[8][4],
                                               const float* __restrict__ zb, int t) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float2 bl = *reinterpret_cast<const float2*>(zb + (2 * ks) * 8 + 2 * t);
    const float2 bh = *reinterpret_cast<const float2*>(zb + (2 * ks + 1) * 8 + 2 * t);
    a[ks][0] = pack_relu(u[2 * ks][0] + bl.x, u[2 * ks][1] + bl.y);
    a[ks][1] = pack_relu(u[2 * ks][2] + bl.x, u[2 * ks][3] + bl.y);
    a[ks][2] = pack_relu(u[2 * ks + 1][0] + bh.x, u[2 * ks + 1][1] + bh.y);
    a[ks][3] = pack_relu(u[2 * ks + 1][2] + bh.x, u[2 * ks + 1][3] + bh.y);
  }
}

__global__ void __launch_bounds__(FT_WARPS * 32, 1)
fcomb_tc_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ mu,
                const float* __restrict__ sigma, const float* __restrict__ eps,
                const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ wmid, const float* __restrict__ bmid,
                const float* __restrict__ wlast, const float* __restrict__ blast,
                float* __restrict__ slice_sums, int N, int L, int C, int nmid, int64_t HW) {
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* W0s = reinterpret_cast<__nv_bfloat16*>(smem);                 // [64][72]
  __nv_bfloat16* Wms = W0s + FT_F * FT_LD;                                     // nmid x [64][72]
  __nv_bfloat16* WLs = Wms + (size_t)nmid * FT_F * FT_LD;                      // [8][72]
  float* bms = reinterpret_cast<float*>(WLs + 8 * FT_LD);                      // nmid x [64]
  float* bls = bms + nmid * FT_F;                                              // [8]
  float* zs = bls + 8;                                                         // [N][16]
  float* zb = zs + (size_t)N * FT_MAXL;                                        // [N][64]

  const int tid = threadIdx.x, b = blockIdx.y;
  // ---- stage weights as bf16, biases, and the per-sample layer-0 bias vectors ----
  for (int i = tid; i < FT_F * FT_F; i += blockDim.x) {
    const int o = i >> 6, k = i & 63;
    W0s[o * FT_LD + k] = __float2bfloat16(__ldg(w0 + (int64_t)o * (FT_F + L) + k));
  }
  for (int m = 0; m < nmid; ++m)
    for (int i = tid; i < FT_F * FT_F; i += blockDim.x) {
      const int o = i >> 6, k = i & 63;
      Wms[(size_t)m * FT_F * FT_LD + o * FT_LD + k] = __float2bfloat16(__ldg(wmid + (int64_t)m * FT_F * FT_F + i));
    }
  for (int i = tid; i < 8 * FT_F; i += blockDim.x) {
    const int o = i >> 6, k = i & 63;
    WLs[o * FT_LD + k] = __float2bfloat16(o < C ? __ldg(wlast + (int64_t)o * FT_F + k) : 0.f);
  }
  for (int i = tid; i < nmid * FT_F; i += blockDim.x) bms[i] = __ldg(bmid + i);
  if (tid < 8) bls[tid] = (tid < C) ? __ldg(blast + tid) : 0.f;
  for (int i = tid; i < N * L; i += blockDim.x) {
    const int n = i / L, l = i % L;
    // z = mu + sigma * eps   (Normal.rsample, probabilistic_unet.py:233)
    zs[n * FT_MAXL + l] = __fadd_rn(__ldg(mu + (int64_t)b * L + l),
                                    __fmul_rn(__ldg(sigma + (int64_t)b * L + l), __ldg(eps + ((int64_t)b * N + n) * L + l)));
  }
  __syncthreads();
  for (int i = tid; i < N * FT_F; i += blockDim.x) {
    const int n = i >> 6, o = i & 63;
    float s = __ldg(b0 + o);
    for (int l = 0; l < L; ++l) s = fmaf(__ldg(w0 + (int64_t)o * (FT_F + L) + FT_F + l), zs[n * FT_MAXL + l], s);
    zb[i] = s;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t block_p0 = (int64_t)blockIdx.x * (FT_WARPS * FT_IT * FT_TP * 16);
  const __nv_bfloat16* fb = feat + (int64_t)b * HW * FT_F;
  const bool c0ok = (2 * t) < C, c1ok = (2 * t + 1) < C;
  const float bl0 = bls[2 * t], bl1 = bls[2 * t + 1];

  for (int it = 0; it < FT_IT; ++it) {
    const int64_t p0 = block_p0 + ((int64_t)it * FT_WARPS + warp) * (FT_TP * 16);
    if (p0 >= HW) break;
    // ---- A fragments of the feature tiles, straight from NHWC global memory ----
    uint32_t a[FT_TP][4][4];
    int64_t pa[FT_TP], pb[FT_TP];
    bool va[FT_TP], vb[FT_TP];
#pragma unroll
    for (int tp = 0; tp < FT_TP; ++tp) {
      pa[tp] = p0 + tp * 16 + g; pb[tp] = pa[tp] + 8;   // the two pixel rows of this thread's fragments
      va[tp] = pa[tp] < HW; vb[tp] = pb[tp] < HW;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int k = ks * 16 + 2 * t;
        a[tp][ks][0] = va[tp] ? __ldg(reinterpret_cast<const uint32_t*>(fb + pa[tp] * FT_F + k)) : 0u;
        a[tp][ks][1] = vb[tp] ? __ldg(reinterpret_cast<const uint32_t*>(fb + pb[tp] * FT_F + k)) : 0u;
        a[tp][ks][2] = va[tp] ? __ldg(reinterpret_cast<const uint32_t*>(fb + pa[tp] * FT_F + k + 8)) : 0u;
        a[tp][ks][3] = vb[tp] ? __ldg(reinterpret_cast<const uint32_t*>(fb + pb[tp] * FT_F + k + 8)) : 0u;
      }
    }
    // ---- shared part of layer 0: u = W0f f (kept in registers for all N samples) ----
    float u[FT_TP][8][4];
    dense64(u, a, W0s, nullptr, g, t);

    float s1[FT_TP][4], s2[FT_TP][4];
#pragma unroll
    for (int tp = 0; tp < FT_TP; ++tp)
#pragma unroll
      for (int j = 0; j < 4; ++j) s1[tp][j] = s2[tp][j] = 0.f;

    for (int n = 0; n < N; ++n) {
      float acc[FT_TP][8][4];
#pragma unroll
      for (int tp = 0; tp < FT_TP; ++tp) bias_relu_pack(a[tp], u[tp], zb + n * FT_F, t);   // h0 = relu(u + zb_n)
      for (int m = 0; m < nmid; ++m) {
        dense64(acc, a, Wms + (size_t)m * FT_F * FT_LD, bms + m * FT_F, g, t);
#pragma unroll
        for (int tp = 0; tp < FT_TP; ++tp) relu_pack(a[tp], acc[tp]);
      }
      // ---- head 64 -> C (N padded to 8), bias rides in as the initial accumulator ----
      float d[FT_TP][4];
      {
        const __nv_bfloat16* wr = WLs + g * FT_LD + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t b0r = *reinterpret_cast<const uint32_t*>(wr + ks * 16);
          const uint32_t b1r = *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8);
#pragma unroll
          for (int tp = 0; tp < FT_TP; ++tp) {
            if (ks == 0) mma_bf16_init(d[tp], a[tp][ks], b0r, b1r, bl0, bl1);
            else mma_bf16_acc(d[tp], a[tp][ks], b0r, b1r);
          }
        }
      }
#pragma unroll
      for (int tp = 0; tp < FT_TP; ++tp) {
        // d[0],d[1]: row g, classes 2t,2t+1;  d[2],d[3]: row g+8
        const float l00 = c0ok ? d[tp][0] : -INFINITY, l01 = c1ok ? d[tp][1] : -INFINITY;
        const float l10 = c0ok ? d[tp][2] : -INFINITY, l11 = c1ok ? d[tp][3] : -INFINITY;
        float m0 = fmaxf(l00, l01), m1 = fmaxf(l10, l11);
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        const float e00 = c0ok ? __expf(l00 - m0) : 0.f, e01 = c1ok ? __expf(l01 - m0) : 0.f;
        const float e10 = c0ok ? __expf(l10 - m1) : 0.f, e11 = c1ok ? __expf(l11 - m1) : 0.f;
        float d0 = e00 + e01, d1 = e10 + e11;
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
        d0 += __shfl_xor_sync(0xffffffffu, d0, 2); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        const float i0 = __fdividef(1.f, d0), i1 = __fdividef(1.f, d1);
        const float p00 = e00 * i0, p01 = e01 * i0, p10 = e10 * i1, p11 = e11 * i1;
        s1[tp][0] += p00; s1[tp][1] += p01; s1[tp][2] += p10; s1[tp][3] += p11;
        s2[tp][0] = fmaf(p00, p00, s2[tp][0]); s2[tp][1] = fmaf(p01, p01, s2[tp][1]);
        s2[tp][2] = fmaf(p10, p10, s2[tp][2]); s2[tp][3] = fmaf(p11, p11, s2[tp][3]);
      }
    }
    // ---- sums: [b][0/1][c][p] ----
    float* o1 = slice_sums + ((int64_t)b * 2 + 0) * C * HW;
    float* o2 = slice_sums + ((int64_t)b * 2 + 1) * C * HW;
#pragma unroll
    for (int tp = 0; tp < FT_TP; ++tp) {
      if (c0ok) {
        if (va[tp]) { o1[(int64_t)(2 * t) * HW + pa[tp]] = s1[tp][0]; o2[(int64_t)(2 * t) * HW + pa[tp]] = s2[tp][0]; }
        if (vb[tp]) { o1[(int64_t)(2 * t) * HW + pb[tp]] = s1[tp][2]; o2[(int64_t)(2 * t) * HW + pb[tp]] = s2[tp][2]; }
      }
      if (c1ok) {
        if (va[tp]) { o1[(int64_t)(2 * t + 1) * HW + pa[tp]] = s1[tp][1]; o2[(int64_t)(2 * t + 1) * HW + pa[tp]] = s2[tp][1]; }
        if (vb[tp]) { o1[(int64_t)(2 * t + 1) * HW + pb[tp]] = s1[tp][3]; o2[(int64_t)(2 * t + 1) * HW + pb[tp]] = s2[tp][3]; }
      }
    }
  }
}

}  // namespace pmu

using namespace pmu;

// mma.sync (legacy tensor path) version: used for no_convs_fcomb > 4 and as an A/B reference
// (PMU_FCOMB_MMA_SYNC=1); the default entry point lives in fcomb_tc6.cu.
extern "C" int pmu_fcomb_softmax_accum_bf16_mma(const void* feat, const float* mu, const float* sigma,
                                            const float* eps, const float* w0, const float* b0,
                                            const float* wmid, const float* bmid, const float* wlast,
                                            const float* blast, float* slice_sums, int B, int N, int L,
                                            int C, int nl, int64_t HW, void* stream) {
  PMU_CHECK_ARG(feat && mu && sigma && eps && w0 && b0 && wlast && blast && slice_sums,
                "pmu_fcomb_softmax_accum_bf16: null pointer");
  PMU_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && HW > 0, "pmu_fcomb_softmax_accum_bf16: bad shape");
  PMU_CHECK_ARG(nl >= 2 && (nl == 2 || (wmid && bmid)), "pmu_fcomb_softmax_accum_bf16: no_convs_fcomb >= 2; mid weights needed for > 2");
  PMU_CHECK_SUPPORTED(L >= 1 && L <= FT_MAXL && C >= 1 && C <= 8, "pmu_fcomb_softmax_accum_bf16: needs L <= 16, C <= 8 (got L=%d C=%d)", L, C);
  PMU_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 3u) == 0, "pmu_fcomb_softmax_accum_bf16: feat must be 4-byte aligned");
  const int nmid = nl - 2;
  const size_t smem = sizeof(__nv_bfloat16) * ((size_t)(1 + nmid) * FT_F * FT_LD + 8 * FT_LD) +
                      sizeof(float) * ((size_t)nmid * FT_F + 8 + (size_t)N * FT_MAXL + (size_t)N * FT_F);
  PMU_CHECK_SUPPORTED(smem <= 200 * 1024, "pmu_fcomb_softmax_accum_bf16: N=%d nl=%d needs %zu B of shared memory", N, nl, smem);
  PMU_CUDA(cudaFuncSetAttribute(fcomb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)cdiv64(HW, FT_WARPS * FT_IT * FT_TP * 16), B);
  fcomb_tc_kernel<<<grid, FT_WARPS * 32, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(feat), mu, sigma, eps, w0, b0, wmid, bmid, wlast, blast,
      slice_sums, N, L, C, nmid, HW);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
