// fp32 NCHW layer ops on CUDA cores — the parity mode of the network (north star:
// "fp32 probabilities within 1e-4 abs").  Same layouts as the reference's torch ops.
//
//   conv3x3 (+folded BN, +ReLU, two-source K loop for the skip concat)  unet_parts.py:15-20,58-66
//   convT 2x2 stride 2 (+F.pad canvas)                                  unet_parts.py:52,61-62
//   MaxPool2d(2) / AvgPool2d(2,2,ceil_mode=True)                        unet_parts.py:33, probabilistic_unet.py:36
//   conv1x1                                                             unet_parts.py:73
//   Gaussian head (mean over H, W + 1x1 conv)                           probabilistic_unet.py:97-108
//   Fcomb over N samples (+ fused softmax / sum / sum-of-squares)       probabilistic_unet.py:155-181
#include "pmu_common.cuh"
#include "h16.cuh"
#include <type_traits>

namespace pmu {

// =====================================================================================
// conv3x3 pad 1.  Block = 256 threads, output tile 8 rows x 32 cols x 32 couts.
// thread: 4 consecutive pixels x 8 couts; K loop over input channels in chunks of 8.
// =====================================================================================
constexpr int CV_TH = 8, CV_TW = 32, CV_CO = 32, CV_CI = 8;
constexpr int CV_ROWSTRIDE = 37;  // == 1 (mod 4): 4-px-strided lanes x 4 rows hit 32 distinct banks

__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                   const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int H, int W, int Cout, int relu, int tiles_x) {
  __shared__ float in_s[CV_CI][CV_TH + 2][CV_ROWSTRIDE];
  __shared__ __align__(16) float w_s[CV_CO / 8][CV_CI * 9][8];

  const int tid = threadIdx.x;
  const int pg = tid & 63, cg = tid >> 6;
  const int py = pg >> 3, px = (pg & 7) * 4;
  const int ty0 = (blockIdx.x / tiles_x) * CV_TH, tx0 = (blockIdx.x % tiles_x) * CV_TW;
  const int co0 = blockIdx.y * CV_CO;
  const int b = blockIdx.z;
  const int Cin = C0 + C1;
  const int64_t HW = (int64_t)H * W;

  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CV_CI) {
    // ---- stage input tile (with halo, zero padded) ----
    for (int idx = tid; idx < CV_CI * (CV_TH + 2) * (CV_TW + 2); idx += 256) {
      const int ci = idx / ((CV_TH + 2) * (CV_TW + 2));
      const int rem = idx % ((CV_TH + 2) * (CV_TW + 2));
      const int r = rem / (CV_TW + 2), c = rem % (CV_TW + 2);
      const int gy = ty0 + r - 1, gx = tx0 + c - 1, gc = ci0 + ci;
      float v = 0.f;
      if (gc < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const float* src = (gc < C0) ? (x0 + ((int64_t)b * C0 + gc) * HW)
                                     : (x1 + ((int64_t)b * C1 + (gc - C0)) * HW);
        v = __ldg(src + (int64_t)gy * W + gx);
      }
      in_s[ci][r][c] = v;
    }
    // ---- stage weights: w[co][Cin][9] -> w_s[co/8][ci*9+tap][co%8] ----
    for (int idx = tid; idx < CV_CO * CV_CI * 9; idx += 256) {
      const int co = idx / (CV_CI * 9), rem = idx % (CV_CI * 9);
      const int ci = rem / 9;
      float v = 0.f;
      if (co0 + co < Cout && ci0 + ci < Cin)
        v = __ldg(w + ((int64_t)(co0 + co) * Cin + ci0) * 9 + rem);
      w_s[co >> 3][rem][co & 7] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CV_CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float iv[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) iv[j] = in_s[ci][py + ky][px + j];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 wa = *reinterpret_cast<const float4*>(&w_s[cg][ci * 9 + ky * 3 + kx][0]);
          const float4 wb = *reinterpret_cast<const float4*>(&w_s[cg][ci * 9 + ky * 3 + kx][4]);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(iv[p + kx], wv[c], acc[p][c]);
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue ----
  const int oy = ty0 + py, ox = tx0 + px;
  if (oy >= H) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + cg * 8 + c;
    if (co >= Cout) continue;
    const float bv = bias ? __ldg(bias + co) : 0.f;
    float* dst = y + ((int64_t)b * Cout + co) * HW + (int64_t)oy * W + ox;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (ox + p < W) {
        float v = acc[p][c] + bv;
        dst[p] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// =====================================================================================
// conv1x1: thread per pixel, 8 couts per pass (grid.y over cout groups).
// =====================================================================================
__global__ void __launch_bounds__(256)
conv1x1_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                   int64_t HW, int relu) {
  const int b = blockIdx.z, co0 = blockIdx.y * 8;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  const float* xp = x + (int64_t)b * Cin * HW + p;
  for (int ci = 0; ci < Cin; ++ci) {
    const float v = __ldg(xp + (int64_t)ci * HW);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (co0 + c < Cout) acc[c] = fmaf(v, __ldg(w + (int64_t)(co0 + c) * Cin + ci), acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (co0 + c < Cout) {
      float v = acc[c] + (bias ? __ldg(bias + co0 + c) : 0.f);
      y[((int64_t)b * Cout + co0 + c) * HW + p] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// =====================================================================================
// ConvTranspose2d k=2 s=2: thread per input pixel, 8 couts x 4 phases per pass.
// =====================================================================================
__global__ void __launch_bounds__(128)
convt2x2_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                    int H, int W, int Ho, int Wo, int padT, int padL) {
  const int b = blockIdx.z, co0 = blockIdx.y * 8;
  const int64_t HW = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (p >= HW) return;
  const int h = (int)(p / W), ww = (int)(p % W);
  float acc[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;
  const float* xp = x + (int64_t)b * Cin * HW + p;
  const bool full = (co0 + 8 <= Cout);
  for (int ci = 0; ci < Cin; ++ci) {
    const float v = __ldg(xp + (int64_t)ci * HW);
    const float* wp = w + ((int64_t)ci * Cout + co0) * 4;  // w[ci][co][i][j]
    if (full) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp) + c);
        acc[c][0] = fmaf(v, wv.x, acc[c][0]); acc[c][1] = fmaf(v, wv.y, acc[c][1]);
        acc[c][2] = fmaf(v, wv.z, acc[c][2]); acc[c][3] = fmaf(v, wv.w, acc[c][3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (co0 + c < Cout)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[c][k] = fmaf(v, __ldg(wp + c * 4 + k), acc[c][k]);
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (co0 + c >= Cout) continue;
    const float bv = bias ? __ldg(bias + co0 + c) : 0.f;
    float* dst = y + ((int64_t)b * Cout + co0 + c) * Ho * Wo;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int oy = padT + 2 * h + i, ox = padL + 2 * ww;
      float2 o = make_float2(acc[c][i * 2 + 0] + bv, acc[c][i * 2 + 1] + bv);
      dst[(int64_t)oy * Wo + ox] = o.x;
      dst[(int64_t)oy * Wo + ox + 1] = o.y;
    }
  }
}

// =====================================================================================
// 2x2 stride-2 pooling
// =====================================================================================
__global__ void __launch_bounds__(256)
pool2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t BC, int H, int W,
                 int Ho, int Wo, int mode) {
  const int64_t total = BC * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const int64_t bc = i / ((int64_t)Wo * Ho);
    const float* src = x + bc * H * W;
    float m = -INFINITY, s = 0.f;
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < H && ix < W) {
          const float v = __ldg(src + (int64_t)iy * W + ix);
          m = fmaxf(m, v);
          s += v;
          ++cnt;
        }
      }
    y[i] = (mode == PMU_POOL_MAX) ? m : s / (float)cnt;
  }
}

// =====================================================================================
// Gaussian head: block per slice; warp per channel for the spatial mean, then one warp
// per output row of the 1x1 conv.
// =====================================================================================
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// NHWC = false: enc[B][C][hw];  NHWC = true: enc[B][hw][C]
template <typename T, bool NHWC>
__global__ void __launch_bounds__(1024)
gauss_head_kernel(const T* __restrict__ enc, const float* __restrict__ w, const float* __restrict__ bvec,
                  float* __restrict__ mu, float* __restrict__ log_sigma, int C, int hw, int L) {
  extern __shared__ float mean_s[];  // [C]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv = 1.f / (float)hw;
  if (!NHWC) {
    for (int c = warp; c < C; c += nw) {
      const T* p = enc + ((int64_t)b * C + c) * hw;
      float s = 0.f;
      for (int i = lane; i < hw; i += 32) s += to_f32<T>(p[i]);
      s = warp_sum(s);
      if (lane == 0) mean_s[c] = s * inv;
    }
  } else if (sizeof(T) == 2 && C % 8 == 0 && blockDim.x % (C / 8) == 0) {
    // bf16 NHWC: thread = (8-channel group, pixel group); 128-bit loads, `pg` pixel groups reduced through smem
    // (deterministic order; a block per slice streams its 0.5 MB at full per-SM bandwidth instead of 2 B per thread)
    const int groups = C / 8, pgs = blockDim.x / groups;
    const int cg = threadIdx.x % groups, pg = threadIdx.x / groups;
    float* part = mean_s + C;          // [pgs][C]
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const uint4* p = reinterpret_cast<const uint4*>(enc + (int64_t)b * hw * C) + cg;
#pragma unroll 4
    for (int i = pg; i < hw; i += pgs) {
      const uint4 v = __ldg(p + (int64_t)i * groups);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f2 = unpack16<sizeof(T) == 2 && !std::is_same<T, __nv_bfloat16>::value>(u[j]);
        acc[2 * j] += f2.x; acc[2 * j + 1] += f2.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[pg * C + cg * 8 + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int g = 0; g < pgs; ++g) s += part[g * C + c];
      mean_s[c] = s * inv;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const T* p = enc + (int64_t)b * hw * C + c;
      float s = 0.f;
      for (int i = 0; i < hw; ++i) s += to_f32<T>(p[(int64_t)i * C]);
      mean_s[c] = s * inv;
    }
  }
  __syncthreads();
  for (int o = warp; o < 2 * L; o += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(mean_s[c], __ldg(w + (int64_t)o * C + c), s);
    s = warp_sum(s);
    if (lane == 0) {
      s += __ldg(bvec + o);
      if (o < L) mu[(int64_t)b * L + o] = s;
      else log_sigma[(int64_t)b * L + (o - L)] = s;
    }
  }
}

// =====================================================================================
// Fcomb, fp32, F <= 64.  Block = P pixels of one slice; activations live in shared memory
// as [F][P]; all layer weights are staged once per block, transposed to [k][o] so that a
// thread reads 4 output weights with one broadcast LDS.128.
//   layer 0 is split: u = W0[:, :F] f + b0 (once per pixel) and zb_n = W0[:, F:] z_n (per sample)
//   — the identity W0 [f; z] + b0 = W0_f f + (W0_z z + b0) of SURVEY.md App. A.
// =====================================================================================
constexpr int FC_P = 128;
constexpr int FC_MAXC = 8;

template <typename TF, bool NHWC>
__global__ void __launch_bounds__(FC_P)
fcomb_f32_kernel(const TF* __restrict__ feat, const float* __restrict__ z, const float* __restrict__ w0,
                 const float* __restrict__ b0, const float* __restrict__ wmid,
                 const float* __restrict__ bmid, const float* __restrict__ wlast,
                 const float* __restrict__ blast, float* __restrict__ logits,
                 float* __restrict__ slice_sums, int N, int F, int L, int C, int nl, int64_t HW) {
  extern __shared__ __align__(16) float smem[];
  const int Fp = (F + 3) & ~3;                 // padded output count (multiple of 4)
  float* u = smem;                             // [F][P]
  float* hA = u + (size_t)F * FC_P;            // [F][P]
  float* hB = hA + (size_t)F * FC_P;           // [F][P]
  float* wS = hB + (size_t)F * FC_P;           // (nl-1) matrices [F][Fp] transposed: wS[m][k][o]
  float* wL = wS + (size_t)(nl - 1) * F * Fp;  // [C][F] last layer
  float* zb = wL + (size_t)FC_MAXC * F;        // [Fp] per-sample bias of layer 0
  const int tid = threadIdx.x, b = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * FC_P + tid;
  const bool live = p < HW;

  // ---- stage weights (transposed) ----
  for (int idx = tid; idx < F * Fp; idx += FC_P) {
    const int k = idx / Fp, o = idx % Fp;
    wS[idx] = (o < F) ? __ldg(w0 + (int64_t)o * (F + L) + k) : 0.f;
  }
  for (int m = 0; m < nl - 2; ++m)
    for (int idx = tid; idx < F * Fp; idx += FC_P) {
      const int k = idx / Fp, o = idx % Fp;
      wS[(size_t)(m + 1) * F * Fp + idx] = (o < F) ? __ldg(wmid + ((int64_t)m * F + o) * F + k) : 0.f;
    }
  for (int idx = tid; idx < C * F; idx += FC_P) wL[idx] = __ldg(wlast + idx);
  // ---- stage features ----
  for (int k = 0; k < F; ++k) {
    float v = 0.f;
    if (live) {
      if (NHWC) v = to_f32<TF>(feat[((int64_t)b * HW + p) * F + k]);
      else v = to_f32<TF>(feat[((int64_t)b * F + k) * HW + p]);
    }
    hA[k * FC_P + tid] = v;
  }
  __syncthreads();
  // ---- shared part of layer 0 ----
  for (int o = 0; o < F; o += 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < F; ++k) {
      const float h = hA[k * FC_P + tid];
      const float4 wv = *reinterpret_cast<const float4*>(wS + (size_t)k * Fp + o);
      a[0] = fmaf(h, wv.x, a[0]); a[1] = fmaf(h, wv.y, a[1]);
      a[2] = fmaf(h, wv.z, a[2]); a[3] = fmaf(h, wv.w, a[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (o + j < F) u[(o + j) * FC_P + tid] = a[j] + __ldg(b0 + o + j);
  }

  float s1[FC_MAXC], s2[FC_MAXC];
#pragma unroll
  for (int c = 0; c < FC_MAXC; ++c) s1[c] = s2[c] = 0.f;

  for (int n = 0; n < N; ++n) {
    __syncthreads();  // previous sample done with zb
    if (tid < F) {
      float s = 0.f;
      for (int l = 0; l < L; ++l)
        s = fmaf(__ldg(w0 + (int64_t)tid * (F + L) + F + l), __ldg(z + ((int64_t)b * N + n) * L + l), s);
      zb[tid] = s;
    }
    __syncthreads();
    float* hin = hB;
    float* hout = hA;
    for (int k = 0; k < F; ++k) hin[k * FC_P + tid] = fmaxf(u[k * FC_P + tid] + zb[k], 0.f);
    for (int m = 0; m < nl - 2; ++m) {
      const float* wm = wS + (size_t)(m + 1) * F * Fp;
      const float* bm = bmid + (int64_t)m * F;
      for (int o = 0; o < F; o += 4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < F; ++k) {
          const float h = hin[k * FC_P + tid];
          const float4 wv = *reinterpret_cast<const float4*>(wm + (size_t)k * Fp + o);
          a[0] = fmaf(h, wv.x, a[0]); a[1] = fmaf(h, wv.y, a[1]);
          a[2] = fmaf(h, wv.z, a[2]); a[3] = fmaf(h, wv.w, a[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o + j < F) hout[(o + j) * FC_P + tid] = fmaxf(a[j] + __ldg(bm + o + j), 0.f);
      }
      float* t = hin; hin = hout; hout = t;
    }
    // ---- last layer + softmax ----
    float lg[FC_MAXC];
#pragma unroll
    for (int c = 0; c < FC_MAXC; ++c) lg[c] = (c < C) ? __ldg(blast + c) : -INFINITY;
    for (int k = 0; k < F; ++k) {
      const float h = hin[k * FC_P + tid];
#pragma unroll
      for (int c = 0; c < FC_MAXC; ++c)
        if (c < C) lg[c] = fmaf(h, wL[c * F + k], lg[c]);
    }
    if (live) {
      if (logits) {
#pragma unroll
        for (int c = 0; c < FC_MAXC; ++c)
          if (c < C) logits[(((int64_t)b * N + n) * C + c) * HW + p] = lg[c];
      }
      float mx = lg[0];
#pragma unroll
      for (int c = 1; c < FC_MAXC; ++c) mx = fmaxf(mx, lg[c]);
      float e[FC_MAXC], den = 0.f;
#pragma unroll
      for (int c = 0; c < FC_MAXC; ++c) {
        e[c] = (c < C) ? expf(lg[c] - mx) : 0.f;
        den += e[c];
      }
      const float inv = 1.f / den;
#pragma unroll
      for (int c = 0; c < FC_MAXC; ++c) {
        const float pr = e[c] * inv;
        s1[c] += pr;
        s2[c] = fmaf(pr, pr, s2[c]);
      }
    }
  }
  if (live && slice_sums) {
#pragma unroll
    for (int c = 0; c < FC_MAXC; ++c)
      if (c < C) {
        slice_sums[(((int64_t)b * 2 + 0) * C + c) * HW + p] = s1[c];
        slice_sums[(((int64_t)b * 2 + 1) * C + c) * HW + p] = s2[c];
      }
  }
}

}  // namespace pmu

using namespace pmu;

extern "C" int pmu_conv3x3_f32(const float* x0, int C0, const float* x1, int C1, const float* w,
                               const float* bias, float* y, int B, int H, int W, int Cout, int relu,
                               void* stream) {
  PMU_CHECK_ARG(x0 && w && y, "pmu_conv3x3_f32: null pointer");
  PMU_CHECK_ARG(C0 > 0 && C1 >= 0 && (C1 == 0 || x1), "pmu_conv3x3_f32: bad channel counts / x1");
  PMU_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cout > 0, "pmu_conv3x3_f32: bad shape");
  PMU_CHECK_ARG(B <= 65535 && cdiv(Cout, CV_CO) <= 65535, "pmu_conv3x3_f32: batch/cout too large");
  const int tiles_x = cdiv(W, CV_TW), tiles_y = cdiv(H, CV_TH);
  dim3 grid(tiles_x * tiles_y, cdiv(Cout, CV_CO), B);
  conv3x3_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, w, bias, y, H, W, Cout, relu, tiles_x);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

// implemented in train_f32.cu: the register-tiled 1x1 kernel (per-batch bias variant; bias_bstride = 0 -> shared bias)
extern "C" int pmu_conv1x1_bb_f32(const float* x, const float* w, int ldw, const float* bias, int bias_bstride, float* y,
                                  int B, int Cin, int Cout, int64_t HW, int relu, void* stream);

extern "C" int pmu_conv1x1_f32(const float* x, const float* w, const float* bias, float* y, int B,
                               int Cin, int Cout, int64_t HW, int relu, void* stream) {
  PMU_CHECK_ARG(x && w && y && B > 0 && Cin > 0 && Cout > 0 && HW > 0, "pmu_conv1x1_f32: bad arguments");
  PMU_CHECK_ARG(B <= 65535, "pmu_conv1x1_f32: batch too large");
  if (Cin <= 1536) return pmu_conv1x1_bb_f32(x, w, Cin, bias, 0, y, B, Cin, Cout, HW, relu, stream);
  dim3 grid((unsigned)cdiv64(HW, 256), cdiv(Cout, 8), B);
  conv1x1_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, Cin, Cout, HW, relu);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_convt2x2_f32(const float* x, const float* w, const float* bias, float* y, int B,
                                int Cin, int Cout, int H, int W, int Ho, int Wo, int padT, int padL,
                                void* stream) {
  PMU_CHECK_ARG(x && w && y && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "pmu_convt2x2_f32: bad arguments");
  PMU_CHECK_ARG(padT >= 0 && padL >= 0 && Ho >= 2 * H + padT && Wo >= 2 * W + padL,
                "pmu_convt2x2_f32: canvas %dx%d too small for %dx%d at (%d,%d)", Ho, Wo, 2 * H, 2 * W, padT, padL);
  PMU_CHECK_ARG(B <= 65535 && aligned16(w), "pmu_convt2x2_f32: batch too large or weights unaligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (Ho != 2 * H || Wo != 2 * W)
    PMU_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)B * Cout * Ho * Wo, st));
  dim3 grid((unsigned)cdiv64((int64_t)H * W, 128), cdiv(Cout, 8), B);
  convt2x2_f32_kernel<<<grid, 128, 0, st>>>(x, w, bias, y, Cin, Cout, H, W, Ho, Wo, padT, padL);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pool2_f32(const float* x, float* y, int B, int C, int H, int W, int mode, void* stream) {
  PMU_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "pmu_pool2_f32: bad arguments");
  PMU_CHECK_ARG(mode == PMU_POOL_MAX || mode == PMU_POOL_AVG_CEIL, "pmu_pool2_f32: unknown mode %d", mode);
  const int Ho = (mode == PMU_POOL_MAX) ? H / 2 : (H + 1) / 2;
  const int Wo = (mode == PMU_POOL_MAX) ? W / 2 : (W + 1) / 2;
  PMU_CHECK_ARG(Ho > 0 && Wo > 0, "pmu_pool2_f32: input too small");
  const int64_t total = (int64_t)B * C * Ho * Wo;
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 16);
  pool2_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, (int64_t)B * C, H, W, Ho, Wo, mode);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_gauss_head_f32(const float* enc, const float* w, const float* b, float* mu,
                                  float* log_sigma, int B, int C, int h, int w_, int L, void* stream) {
  PMU_CHECK_ARG(enc && w && b && mu && log_sigma, "pmu_gauss_head_f32: null pointer");
  PMU_CHECK_ARG(B > 0 && C > 0 && h > 0 && w_ > 0 && L > 0 && C <= 12288, "pmu_gauss_head_f32: bad shape");
  gauss_head_kernel<float, false><<<B, 256, C * sizeof(float), (cudaStream_t)stream>>>(enc, w, b, mu, log_sigma, C, h * w_, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_gauss_head_bf16(const void* enc, const float* w, const float* b, float* mu,
                                   float* log_sigma, int B, int C, int h, int w_, int L, int f16, void* stream) {
  PMU_CHECK_ARG(enc && w && b && mu && log_sigma, "pmu_gauss_head_bf16: null pointer");
  PMU_CHECK_ARG(B > 0 && C > 0 && h > 0 && w_ > 0 && L > 0 && C <= 12288, "pmu_gauss_head_bf16: bad shape");
  // vector path: threads = (C/8 channel groups) x (pixel groups), as many pixel groups as fit 1024 threads
  PMU_CHECK_ARG(aligned16(enc), "pmu_gauss_head_bf16: enc must be 16-byte aligned");
  int threads = 256;
  size_t smem = C * sizeof(float);
  if (C % 8 == 0) {
    const int groups = C / 8;
    if (groups <= 1024) {
      int pgs = std::max(1, std::min(1024 / groups, h * w_));
      while (pgs > 1 && (groups * pgs) % 32) --pgs;
      if ((groups * pgs) % 32 == 0) threads = groups * pgs;
    }
    if (threads % groups == 0) smem = (size_t)(1 + threads / groups) * C * sizeof(float);   // the kernel takes the vector path
  }
  if (f16) {
    auto kern = gauss_head_kernel<__half, true>;
    if (smem > 48 * 1024) PMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, threads, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(enc), w, b, mu, log_sigma, C, h * w_, L);
  } else {
    auto kern = gauss_head_kernel<__nv_bfloat16, true>;
    if (smem > 48 * 1024) PMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, threads, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(enc), w, b, mu, log_sigma, C, h * w_, L);
  }
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

namespace pmu {
template <typename TF, bool NHWC>
int launch_fcomb_f32(const TF* feat, const float* z, const float* w0, const float* b0, const float* wmid,
                     const float* bmid, const float* wlast, const float* blast, float* logits,
                     float* slice_sums, int B, int N, int F, int L, int C, int nl, int64_t HW,
                     cudaStream_t st) {
  const int Fp = (F + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)3 * F * FC_P + (size_t)(nl - 1) * F * Fp + (size_t)FC_MAXC * F + Fp);
  PMU_CHECK_SUPPORTED(smem <= 227 * 1024, "pmu_fcomb_f32: F=%d nl=%d needs %zu B of shared memory", F, nl, smem);
  auto kern = fcomb_f32_kernel<TF, NHWC>;
  PMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)cdiv64(HW, FC_P), B);
  kern<<<grid, FC_P, smem, st>>>(feat, z, w0, b0, wmid, bmid, wlast, blast, logits, slice_sums, N, F, L, C, nl, HW);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
}  // namespace pmu

extern "C" int pmu_fcomb_f32(const float* feat, const float* z, const float* w0, const float* b0,
                             const float* wmid, const float* bmid, const float* wlast, const float* blast,
                             float* logits, float* slice_sums, int B, int N, int F, int L, int C, int nl,
                             int64_t HW, void* stream) {
  PMU_CHECK_ARG(feat && z && w0 && b0 && wlast && blast, "pmu_fcomb_f32: null pointer");
  PMU_CHECK_ARG(logits || slice_sums, "pmu_fcomb_f32: need logits and/or slice_sums output");
  PMU_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && L > 0 && HW > 0, "pmu_fcomb_f32: bad shape");
  PMU_CHECK_ARG(nl >= 2 && (nl == 2 || (wmid && bmid)), "pmu_fcomb_f32: no_convs_fcomb must be >= 2 (mid weights needed for > 2)");
  PMU_CHECK_SUPPORTED(F > 0 && F <= 64 && C > 0 && C <= FC_MAXC, "pmu_fcomb_f32: supports F <= 64, C <= 8 (got F=%d C=%d)", F, C);
  return launch_fcomb_f32<float, false>(feat, z, w0, b0, wmid, bmid, wlast, blast, logits, slice_sums,
                                        B, N, F, L, C, nl, HW, (cudaStream_t)stream);
}
