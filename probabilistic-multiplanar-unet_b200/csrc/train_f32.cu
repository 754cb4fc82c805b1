// Training step, fp32 NCHW parity mode: train-mode BatchNorm forward /
// backward, convolution weight gradients, pooling / transposed-convolution backward, the
// Gaussian-head, fcomb, cross-entropy and KL backward passes.
//
// Replaces what autograd does for the reference's training step (train.py:85-110):
//   nn.BatchNorm2d in training mode          unet_parts.py:16,19; probabilistic_unet.py:39,44
//   nn.Conv2d / nn.ConvTranspose2d backward  unet_parts.py:15,18,52; probabilistic_unet.py:38,43,137-146
//   nn.MaxPool2d / nn.AvgPool2d backward     unet_parts.py:33; probabilistic_unet.py:36
//   torch.mean + 1x1 conv head backward      probabilistic_unet.py:97-108
//   CrossEntropyLoss / kl.kl_divergence      probabilistic_unet.py:272,288-304
// Data gradients of the 3x3 / 1x1 convolutions reuse the forward kernels (pmu_conv3x3_f32 /
// pmu_conv1x1_f32) with transposed + flipped weights packed by the caller.
//
// Reductions over (batch, pixels) accumulate block partials (fp32) into fp64 with atomics, so
// mean / variance / gradient sums do not lose bits to cancellation or to the summation order.
#include "pmu_common.cuh"

namespace pmu {

// ------------------------------------------------------------------------------------
// block-wide sum of two values (256 threads)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum2(float& a, float& b) {
  __shared__ float sa[8], sb[8];
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = (l < (blockDim.x >> 5)) ? sa[l] : 0.f;
    b = (l < (blockDim.x >> 5)) ? sb[l] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
  }
}

// ------------------------------------------------------------------------------------
// BatchNorm2d training forward / backward.  Every kernel walks rows (b, c) of HW contiguous floats, so there is no
// integer division in the inner loops, and with HW % 4 == 0 (VEC) every access is a 128-bit load / store.
// Per-channel reductions: grid (chunks over HW, C), a block loops over the batch; block partials (fp32) are added to
// fp64 accumulators acc[C][2] with atomics.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <bool VEC>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ y, int B, int C, int64_t HW, int64_t chunk, double* __restrict__ acc) {
  const int c = blockIdx.y;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < HW) ? lo + chunk : HW;
  float s = 0.f, ss = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* row = y + ((int64_t)b * C + c) * HW;
    if (VEC) {
#pragma unroll 4
      for (int64_t p = lo + threadIdx.x * 4; p < hi; p += 1024) {
        const float4 v = ld4(row + p);
        s += (v.x + v.y) + (v.z + v.w);
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
      }
    } else {
      for (int64_t p = lo + threadIdx.x; p < hi; p += 256) {
        const float v = __ldg(row + p);
        s += v; ss = fmaf(v, v, ss);
      }
    }
  }
  block_sum2(s, ss);
  if (threadIdx.x == 0) { atomicAdd(acc + 2 * c, (double)s); atomicAdd(acc + 2 * c + 1, (double)ss); }
}
// mean, biased variance (what normalises the batch); running stats with the unbiased variance
// (torch: running = (1 - m) * running + m * stat).
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double n, float* __restrict__ mean,
                                   float* __restrict__ var, float* __restrict__ run_mean, float* __restrict__ run_var,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[2 * c] / n;
  double v = acc[2 * c + 1] / n - m * m;
  if (v < 0) v = 0;
  mean[c] = (float)m;
  var[c] = (float)v;
  if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
  if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(n > 1 ? v * n / (n - 1) : v);
}
// a = [relu](gamma * (y - mean) * rsqrt(var + eps) + beta); grid (B*C, chunks over HW)
template <bool VEC>
__global__ void __launch_bounds__(256)
bn_act_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ var,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
              float* __restrict__ a, int C, int64_t HW) {
  const int bc = blockIdx.x, c = bc % C;
  const float inv = 1.f / sqrtf(__ldg(var + c) + eps);
  const float m = __ldg(mean + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  const float* src = y + (int64_t)bc * HW;
  float* dst = a + (int64_t)bc * HW;
  auto f = [&](float x) { const float v = (x - m) * inv * g + bt; return relu ? fmaxf(v, 0.f) : v; };
  if (VEC) {
    for (int64_t p = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 4; p < HW; p += (int64_t)gridDim.y * 1024) {
      const float4 v = ld4(src + p);
      *reinterpret_cast<float4*>(dst + p) = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.y * 256 + threadIdx.x; p < HW; p += (int64_t)gridDim.y * 256) dst[p] = f(src[p]);
  }
}
// backward, pass 1: dz = da * (z > 0) with z recomputed from y; acc[c] += {sum dz, sum dz * xhat}
template <bool VEC>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, int relu, int B, int C, int64_t HW, int64_t chunk, double* __restrict__ acc) {
  const int c = blockIdx.y;
  const float inv = 1.f / sqrtf(__ldg(var + c) + eps);
  const float m = __ldg(mean + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < HW) ? lo + chunk : HW;
  float s1 = 0.f, s2 = 0.f;
  auto f = [&](float yv, float d) {
    const float xh = (yv - m) * inv;
    if (relu && !(xh * g + bt > 0.f)) d = 0.f;
    s1 += d; s2 = fmaf(d, xh, s2);
  };
  for (int b = 0; b < B; ++b) {
    const int64_t base = ((int64_t)b * C + c) * HW;
    if (VEC) {
#pragma unroll 2
      for (int64_t p = lo + threadIdx.x * 4; p < hi; p += 1024) {
        const float4 yv = ld4(y + base + p), d = ld4(da + base + p);
        f(yv.x, d.x); f(yv.y, d.y); f(yv.z, d.z); f(yv.w, d.w);
      }
    } else {
      for (int64_t p = lo + threadIdx.x; p < hi; p += 256) f(__ldg(y + base + p), __ldg(da + base + p));
    }
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) { atomicAdd(acc + 2 * c, (double)s1); atomicAdd(acc + 2 * c + 1, (double)s2); }
}
// backward, pass 2: dy = gamma * inv * (dz - mean(dz) - xhat * mean(dz * xhat)); also writes
// dgamma = sum dz * xhat, dbeta = sum dz (block (0, c-th row of b == 0) does it).  bias_acc (nullable, fp64 [C],
// zero-filled): += sum of the dy values this block wrote — the bias gradient of the convolution in front of the
// BatchNorm (mathematically zero; autograd's value is the rounding residue of exactly this sum).
template <bool VEC>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, int relu, const double* __restrict__ acc, double n, float* __restrict__ dy,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, double* __restrict__ bias_acc, int C,
                    int64_t HW) {
  const int bc = blockIdx.x, c = bc % C;
  const float inv = 1.f / sqrtf(__ldg(var + c) + eps);
  const float m = __ldg(mean + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  const float m1 = (float)(acc[2 * c] / n), m2 = (float)(acc[2 * c + 1] / n);
  if (bc == c && blockIdx.y == 0 && threadIdx.x == 0) {
    if (dbeta) dbeta[c] = (float)acc[2 * c];
    if (dgamma) dgamma[c] = (float)acc[2 * c + 1];
  }
  const int64_t base = (int64_t)bc * HW;
  float sb = 0.f, zero = 0.f;
  auto f = [&](float yv, float d) {
    const float xh = (yv - m) * inv;
    if (relu && !(xh * g + bt > 0.f)) d = 0.f;
    const float r = g * inv * (d - m1 - xh * m2);
    sb += r;
    return r;
  };
  if (VEC) {
    for (int64_t p = ((int64_t)blockIdx.y * 256 + threadIdx.x) * 4; p < HW; p += (int64_t)gridDim.y * 1024) {
      const float4 yv = ld4(y + base + p), d = ld4(da + base + p);
      *reinterpret_cast<float4*>(dy + base + p) = make_float4(f(yv.x, d.x), f(yv.y, d.y), f(yv.z, d.z), f(yv.w, d.w));
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.y * 256 + threadIdx.x; p < HW; p += (int64_t)gridDim.y * 256)
      dy[base + p] = f(y[base + p], da[base + p]);
  }
  if (bias_acc) {
    block_sum2(sb, zero);
    if (threadIdx.x == 0) atomicAdd(bias_acc + c, (double)sb);
  }
}

// ------------------------------------------------------------------------------------
// per-channel sums over (B, HW): conv bias gradients.  acc fp64 [C][2] (slot 0) or a plain fp64 [C] (stride 1)
// ------------------------------------------------------------------------------------
__global__ void channel_sum_finalize_kernel(const double* __restrict__ acc, int stride, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[stride * c];
}
// per-row sums: out[r] = sum_p x[r][p]   (rows = B*C); one block per row
__global__ void __launch_bounds__(256)
row_sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  const float* src = x + (int64_t)blockIdx.x * n;
  float s = 0.f, z = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += __ldg(src + i);
  block_sum2(s, z);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------
// conv3x3 weight gradient: dw[co][ci][ky][kx] += sum_{b,h,w} dy[b,co,h,w] * x[b,ci,h+ky-1,w+kx-1]
// Block = 256 threads = 16 (ci pairs) x 16 (co pairs): tile 32 co x 32 ci, every thread 2 x 2 x 9
// accumulators; spatial tiles of 4 rows x 32 columns staged in shared memory (x with halo); the
// three kx taps slide over registers.  Split over (batch, spatial tiles) across blockIdx.y;
// partials are added to dw with fp32 atomics (dw must be zero-filled or hold a running sum).
// ------------------------------------------------------------------------------------
constexpr int WG_CO = 32, WG_CI = 32, WG_TH = 4, WG_TW = 32;
constexpr int WG_XROW = WG_TW + 3;                 // 35 (34 used)
constexpr int WG_XCH = (WG_TH + 2) * WG_XROW + 1;  // 211: odd channel stride -> conflict-free across ci
constexpr int WG_DCH = WG_TH * WG_TW + 1;          // 129

__global__ void __launch_bounds__(256)
conv3x3_wgrad_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                     const float* __restrict__ dy, float* __restrict__ dw, int B, int H, int W, int Cout,
                     int tiles_x, int tiles_y, int units_per_block) {
  __shared__ float x_s[WG_CI * WG_XCH];
  __shared__ float d_s[WG_CO * WG_DCH];
  const int Cin = C0 + C1;
  const int ci_tiles = (Cin + WG_CI - 1) / WG_CI;
  const int co0 = (blockIdx.x / ci_tiles) * WG_CO, ci0 = (blockIdx.x % ci_tiles) * WG_CI;
  const int tid = threadIdx.x, tci = tid & 15, tco = tid >> 4;
  const int64_t HW = (int64_t)H * W;
  const int units = B * tiles_x * tiles_y;
  const int u_lo = blockIdx.y * units_per_block, u_hi = min(units, u_lo + units_per_block);

  float acc[2][2][9];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[a][b][t] = 0.f;

  for (int u = u_lo; u < u_hi; ++u) {
    const int b = u / (tiles_x * tiles_y), r = u % (tiles_x * tiles_y);
    const int ty0 = (r / tiles_x) * WG_TH, tx0 = (r % tiles_x) * WG_TW;
    __syncthreads();
    // stage x tile with halo (zero padded)
    for (int idx = tid; idx < WG_CI * (WG_TH + 2) * (WG_TW + 2); idx += 256) {
      const int ci = idx / ((WG_TH + 2) * (WG_TW + 2)), rem = idx % ((WG_TH + 2) * (WG_TW + 2));
      const int rr = rem / (WG_TW + 2), cc = rem % (WG_TW + 2);
      const int gy = ty0 + rr - 1, gx = tx0 + cc - 1, gc = ci0 + ci;
      float v = 0.f;
      if (gc < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const float* src = (gc < C0) ? (x0 + ((int64_t)b * C0 + gc) * HW) : (x1 + ((int64_t)b * C1 + (gc - C0)) * HW);
        v = __ldg(src + (int64_t)gy * W + gx);
      }
      x_s[ci * WG_XCH + rr * WG_XROW + cc] = v;
    }
    // stage dy tile
    for (int idx = tid; idx < WG_CO * WG_TH * WG_TW; idx += 256) {
      const int co = idx / (WG_TH * WG_TW), rem = idx % (WG_TH * WG_TW);
      const int rr = rem / WG_TW, cc = rem % WG_TW;
      const int gy = ty0 + rr, gx = tx0 + cc, gc = co0 + co;
      float v = 0.f;
      if (gc < Cout && gy < H && gx < W) v = __ldg(dy + ((int64_t)b * Cout + gc) * HW + (int64_t)gy * W + gx);
      d_s[co * WG_DCH + rem] = v;
    }
    __syncthreads();
    const float* xa = x_s + (2 * tci) * WG_XCH;
    const float* xb = xa + WG_XCH;
    const float* da_ = d_s + (2 * tco) * WG_DCH;
    const float* db_ = da_ + WG_DCH;
#pragma unroll
    for (int rr = 0; rr < WG_TH; ++rr) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* ra = xa + (rr + ky) * WG_XROW;
        const float* rb = xb + (rr + ky) * WG_XROW;
        float a0 = ra[0], a1 = ra[1], b0 = rb[0], b1 = rb[1];
#pragma unroll 8
        for (int cc = 0; cc < WG_TW; ++cc) {
          const float a2 = ra[cc + 2], b2 = rb[cc + 2];
          const float d0 = da_[rr * WG_TW + cc], d1 = db_[rr * WG_TW + cc];
          acc[0][0][ky * 3 + 0] = fmaf(d0, a0, acc[0][0][ky * 3 + 0]);
          acc[0][0][ky * 3 + 1] = fmaf(d0, a1, acc[0][0][ky * 3 + 1]);
          acc[0][0][ky * 3 + 2] = fmaf(d0, a2, acc[0][0][ky * 3 + 2]);
          acc[0][1][ky * 3 + 0] = fmaf(d0, b0, acc[0][1][ky * 3 + 0]);
          acc[0][1][ky * 3 + 1] = fmaf(d0, b1, acc[0][1][ky * 3 + 1]);
          acc[0][1][ky * 3 + 2] = fmaf(d0, b2, acc[0][1][ky * 3 + 2]);
          acc[1][0][ky * 3 + 0] = fmaf(d1, a0, acc[1][0][ky * 3 + 0]);
          acc[1][0][ky * 3 + 1] = fmaf(d1, a1, acc[1][0][ky * 3 + 1]);
          acc[1][0][ky * 3 + 2] = fmaf(d1, a2, acc[1][0][ky * 3 + 2]);
          acc[1][1][ky * 3 + 0] = fmaf(d1, b0, acc[1][1][ky * 3 + 0]);
          acc[1][1][ky * 3 + 1] = fmaf(d1, b1, acc[1][1][ky * 3 + 1]);
          acc[1][1][ky * 3 + 2] = fmaf(d1, b2, acc[1][1][ky * 3 + 2]);
          a0 = a1; a1 = a2; b0 = b1; b1 = b2;
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int co = co0 + 2 * tco + a;
    if (co >= Cout) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ci = ci0 + 2 * tci + b;
      if (ci >= Cin) continue;
      float* dst = dw + ((int64_t)co * Cin + ci) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[a][b][t]);
    }
  }
}

// ------------------------------------------------------------------------------------
// conv1x1 weight gradient (kernel further down): dw[co][ci] += sum_{b,p} dy[b,co,p] * x[b,ci,p].  Tile 32 x 32,
// 2 x 2 per thread, 128-pixel chunks in shared memory; split over (batch, chunks) across blockIdx.y.
// ------------------------------------------------------------------------------------
constexpr int W1_P = 128;

// ------------------------------------------------------------------------------------
// 2x2 pooling backward.  MAX: the gradient goes to the first maximum of the window in row-major
// order (torch); AVG_CEIL: dy / (number of in-bounds taps).  Thread per output window.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int H, int W,
                 int Ho, int Wo, int mode, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
  const int64_t bc = i / ((int64_t)Wo * Ho);
  const float g = dy[i];
  const int y0 = 2 * oy, x0 = 2 * ox;
  float* dst = dx + bc * H * W;
  if (mode == PMU_POOL_MAX) {
    const float* src = x + bc * H * W;
    int best = 0;
    float bv = src[(int64_t)y0 * W + x0];
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      const float v = src[(int64_t)(y0 + (t >> 1)) * W + x0 + (t & 1)];
      if (v > bv) { bv = v; best = t; }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[(int64_t)(y0 + (t >> 1)) * W + x0 + (t & 1)] = (t == best) ? g : 0.f;
  } else {
    const int ny = (y0 + 1 < H) ? 2 : 1, nx = (x0 + 1 < W) ? 2 : 1;
    const float v = g / (float)(ny * nx);
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) dst[(int64_t)(y0 + a) * W + x0 + b] = v;
  }
}

// ------------------------------------------------------------------------------------
// ConvTranspose2d k2 s2 backward.
//   dgrad: dx[b,ci,h,w] = sum_{co,i,j} dy[b,co,2h+i,2w+j] * w[ci,co,i,j]   (thread per pixel, 8 ci per pass)
//   wgrad: dw[ci,co,i,j] += sum_{b,h,w} x[b,ci,h,w] * dy[b,co,2h+i,2w+j]   (tile 32 ci x 16 co x 4 phases)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
convt2x2_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int Cin,
                      int Cout, int H, int W) {
  const int b = blockIdx.z, ci0 = blockIdx.y * 8;
  const int64_t HW = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (p >= HW) return;
  const int h = (int)(p / W), x = (int)(p % W);
  const int W2 = 2 * W;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  const float* dyb = dy + (int64_t)b * Cout * 4 * HW + (int64_t)(2 * h) * W2 + 2 * x;
  for (int co = 0; co < Cout; ++co) {
    const float* q = dyb + (int64_t)co * 4 * HW;
    const float2 r0 = *reinterpret_cast<const float2*>(q);
    const float2 r1 = *reinterpret_cast<const float2*>(q + W2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (ci0 + c < Cin) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((int64_t)(ci0 + c) * Cout + co) * 4));
        acc[c] = fmaf(r0.x, wv.x, fmaf(r0.y, wv.y, fmaf(r1.x, wv.z, fmaf(r1.y, wv.w, acc[c]))));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (ci0 + c < Cin) dx[((int64_t)b * Cin + ci0 + c) * HW + p] = acc[c];
}

constexpr int CT_P = 64;
__global__ void __launch_bounds__(256)
convt2x2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int B,
                      int Cin, int Cout, int H, int W, int chunks_per_img, int units_per_block) {
  __shared__ float x_s[32 * (CT_P + 1)];
  __shared__ float d_s[16 * (4 * CT_P + 1)];
  const int co_tiles = (Cout + 15) / 16;
  const int ci0 = (blockIdx.x / co_tiles) * 32, co0 = (blockIdx.x % co_tiles) * 16;
  const int tid = threadIdx.x, tci = tid & 15, tco = tid >> 4;     // 2 ci x 1 co x 4 phases per thread
  const int64_t HW = (int64_t)H * W;
  const int W2 = 2 * W;
  const int units = B * chunks_per_img;
  const int u_lo = blockIdx.y * units_per_block, u_hi = min(units, u_lo + units_per_block);
  float acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[a][t] = 0.f;
  for (int u = u_lo; u < u_hi; ++u) {
    const int b = u / chunks_per_img;
    const int64_t p0 = (int64_t)(u % chunks_per_img) * CT_P;
    __syncthreads();
    for (int idx = tid; idx < 32 * CT_P; idx += 256) {
      const int c = idx / CT_P, p = idx % CT_P;
      float v = 0.f;
      if (p0 + p < HW && ci0 + c < Cin) v = __ldg(x + ((int64_t)b * Cin + ci0 + c) * HW + p0 + p);
      x_s[c * (CT_P + 1) + p] = v;
    }
    for (int idx = tid; idx < 16 * 4 * CT_P; idx += 256) {
      const int c = idx / (4 * CT_P), rem = idx % (4 * CT_P);
      const int t = rem / CT_P, p = rem % CT_P;
      float v = 0.f;
      if (p0 + p < HW && co0 + c < Cout) {
        const int h = (int)((p0 + p) / W), xx = (int)((p0 + p) % W);
        v = __ldg(dy + ((int64_t)b * Cout + co0 + c) * 4 * HW + (int64_t)(2 * h + (t >> 1)) * W2 + 2 * xx + (t & 1));
      }
      d_s[c * (4 * CT_P + 1) + rem] = v;
    }
    __syncthreads();
    const float* xa = x_s + (2 * tci) * (CT_P + 1);
    const float* dd = d_s + tco * (4 * CT_P + 1);
#pragma unroll 8
    for (int p = 0; p < CT_P; ++p) {
      const float a0 = xa[p], a1 = xa[CT_P + 1 + p];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float d = dd[t * CT_P + p];
        acc[0][t] = fmaf(a0, d, acc[0][t]);
        acc[1][t] = fmaf(a1, d, acc[1][t]);
      }
    }
  }
  const int co = co0 + tco;
  if (co < Cout) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ci = ci0 + 2 * tci + a;
      if (ci < Cin)
#pragma unroll
        for (int t = 0; t < 4; ++t) atomicAdd(dw + ((int64_t)ci * Cout + co) * 4 + t, acc[a][t]);
    }
  }
}

// ------------------------------------------------------------------------------------
// elementwise pieces
// ------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(const float* __restrict__ a, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = (a[i] > 0.f) ? dy[i] : 0.f;
}
__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] += src[i];
}
// dlogits[b,c,p] = scale * (softmax_c(logits[b,:,p]) - [c == target[b,p]])   (CrossEntropyLoss, reduction sum)
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, float scale, float* __restrict__ dl,
              int C, int64_t HW) {
  const int b = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const float* lp = logits + (int64_t)b * C * HW + p;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, lp[(int64_t)c * HW]);
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(lp[(int64_t)c * HW] - mx);
  const int t = (int)target[(int64_t)b * HW + p];
  float* dp = dl + (int64_t)b * C * HW + p;
  for (int c = 0; c < C; ++c) dp[(int64_t)c * HW] = scale * (expf(lp[(int64_t)c * HW] - mx) / den - (c == t ? 1.f : 0.f));
}
// analytic KL(q||p) backward (kl.kl_divergence of Independent Normals, probabilistic_unet.py:272), scaled
__global__ void kl_bwd_kernel(const float* __restrict__ mu_q, const float* __restrict__ ls_q, const float* __restrict__ mu_p,
                              const float* __restrict__ ls_p, float scale, float* __restrict__ dmu_q, float* __restrict__ dls_q,
                              float* __restrict__ dmu_p, float* __restrict__ dls_p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float vq = expf(2.f * ls_q[i]), vp = expf(2.f * ls_p[i]), d = mu_q[i] - mu_p[i];
  dmu_q[i] = scale * d / vp;
  dmu_p[i] = -scale * d / vp;
  dls_q[i] = scale * (vq / vp - 1.f);
  dls_p[i] = scale * (1.f - (vq + d * d) / vp);
}

// ------------------------------------------------------------------------------------
// Gaussian head backward (probabilistic_unet.py:97-108).  Block per batch item.
//   pooled[c] = mean_hw enc[b,c]; out[j] = w[j,:] . pooled + bias[j]; dout = [dmu | dlog_sigma]
//   denc[b,c,:] = (sum_j dout[j] w[j,c]) / hw;  dw[j,c] += dout[j] pooled[c];  db[j] += dout[j]
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gauss_head_bwd_kernel(const float* __restrict__ enc, const float* __restrict__ w, const float* __restrict__ dmu,
                      const float* __restrict__ dls, float* __restrict__ denc, float* __restrict__ dw,
                      float* __restrict__ db, int C, int hw, int L) {
  const int b = blockIdx.x;
  extern __shared__ float dout[];       // [2L]
  for (int j = threadIdx.x; j < 2 * L; j += 256) dout[j] = (j < L) ? dmu[b * L + j] : dls[b * L + j - L];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const float* e = enc + ((int64_t)b * C + c) * hw;
    float s = 0.f;
    for (int p = 0; p < hw; ++p) s += e[p];
    const float pooled = s / (float)hw;
    float dp = 0.f;
    for (int j = 0; j < 2 * L; ++j) {
      dp = fmaf(dout[j], __ldg(w + (int64_t)j * C + c), dp);
      atomicAdd(dw + (int64_t)j * C + c, dout[j] * pooled);
    }
    dp /= (float)hw;
    float* d = denc + ((int64_t)b * C + c) * hw;
    for (int p = 0; p < hw; ++p) d[p] = dp;
  }
  if (threadIdx.x < 2 * L) atomicAdd(db + threadIdx.x, dout[threadIdx.x]);
}

// ------------------------------------------------------------------------------------
// fcomb, training form (probabilistic_unet.py:155-181): layer 0 split into the feature part
// and the per-slice vector zb[b,co] = b0[co] + sum_l W0[co, F+l] z[b,l]; 1x1 conv with a bias per
// (batch item, output channel).
// ------------------------------------------------------------------------------------
__global__ void fcomb_zbias_kernel(const float* __restrict__ z, const float* __restrict__ w0, const float* __restrict__ b0,
                                   float* __restrict__ zb, int B, int F, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * F) return;
  const int b = i / F, co = i % F;
  float s = b0[co];
  for (int l = 0; l < L; ++l) s = fmaf(w0[(int64_t)co * (F + L) + F + l], z[b * L + l], s);
  zb[i] = s;
}
// given rs[b,co] = sum_p dh0[b,co,p]:  dz[b,l] = sum_co W0[co,F+l] rs[b,co];
// dW0[co,F+l] += sum_b z[b,l] rs[b,co];  db0[co] += sum_b rs[b,co].   One block.
__global__ void fcomb_zbias_bwd_kernel(const float* __restrict__ rs, const float* __restrict__ z, const float* __restrict__ w0,
                                       float* __restrict__ dz, float* __restrict__ dw0, float* __restrict__ db0, int B, int F, int L) {
  for (int i = threadIdx.x; i < B * L; i += blockDim.x) {
    const int b = i / L, l = i % L;
    float s = 0.f;
    for (int co = 0; co < F; ++co) s = fmaf(w0[(int64_t)co * (F + L) + F + l], rs[b * F + co], s);
    dz[i] = s;
  }
  for (int i = threadIdx.x; i < F * L; i += blockDim.x) {
    const int co = i / L, l = i % L;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(z[b * L + l], rs[b * F + co], s);
    dw0[(int64_t)co * (F + L) + F + l] += s;
  }
  for (int co = threadIdx.x; co < F; co += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += rs[b * F + co];
    db0[co] += s;
  }
}
// y[b,co,p] = [relu](sum_ci w[co*ldw + ci] x[b,ci,p] + bias[b*bias_bstride + co]).
// Register tile: 4 consecutive pixels x 8 couts per thread.  The 8 x Cin weight slab of the block sits in shared memory
// as [ci][8], so one ci step is ONE 128-bit global load (4 pixels) + TWO 128-bit broadcast shared loads for 32 FMAs
// (the thread-per-pixel version issued 9 loads per 8 FMAs and was bound by the load/store unit).  The sum over ci keeps
// its sequential fmaf order, so results are bit-identical with the previous kernel.
template <bool VEC>
__global__ void __launch_bounds__(256)
conv1x1_bb_kernel(const float* __restrict__ x, const float* __restrict__ w, int ldw, const float* __restrict__ bias,
                  int bias_bstride, float* __restrict__ y, int Cin, int Cout, int64_t HW, int relu) {
  extern __shared__ float ws[];                       // [Cin][8]
  const int b = blockIdx.z, co0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < Cin * 8; i += 256) {
    const int ci = i >> 3, c = i & 7;
    ws[i] = (co0 + c < Cout) ? __ldg(w + (int64_t)(co0 + c) * ldw + ci) : 0.f;
  }
  __syncthreads();
  const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (p >= HW) return;
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
  const float* xp = x + (int64_t)b * Cin * HW + p;
  const int npx = (HW - p < 4) ? (int)(HW - p) : 4;
  for (int ci = 0; ci < Cin; ++ci) {
    float v[4];
    if (VEC) {
      const float4 t = *reinterpret_cast<const float4*>(xp + (int64_t)ci * HW);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (q < npx) ? __ldg(xp + (int64_t)ci * HW + q) : 0.f;
    }
    const float4 wa = *reinterpret_cast<const float4*>(ws + ci * 8), wb = *reinterpret_cast<const float4*>(ws + ci * 8 + 4);
    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = fmaf(v[q], wv[c], acc[q][c]);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (co0 + c >= Cout) continue;
    const float bv = bias ? __ldg(bias + (int64_t)b * bias_bstride + co0 + c) : 0.f;
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { o[q] = acc[q][c] + bv; if (relu) o[q] = fmaxf(o[q], 0.f); }
    float* yp = y + ((int64_t)b * Cout + co0 + c) * HW + p;
    if (VEC) *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
    else
#pragma unroll
      for (int q = 0; q < 4; ++q) if (q < npx) yp[q] = o[q];
  }
}

static inline unsigned ew_grid(int64_t n) { return (unsigned)std::min<int64_t>(cdiv64(n, 256), (int64_t)sm_count() * 16); }

}  // namespace pmu

using namespace pmu;

// split HW into chunks (multiples of 1024 = 256 threads x float4) so that chunks * C blocks fill the machine a few times over
static void stat_chunks(int C, int64_t HW, int64_t* chunk, int* nchunks) {
  int64_t want = std::max<int64_t>(1, (8ll * sm_count() + C - 1) / C);
  int64_t ch = std::max<int64_t>(1024, cdiv64(HW, want));
  ch = ((ch + 1023) / 1024) * 1024;
  *chunk = ch;
  *nchunks = (int)cdiv64(HW, ch);
}
static inline bool rows_vec(int64_t HW, const void* a, const void* b = nullptr, const void* c = nullptr) {
  return HW % 4 == 0 && aligned16(a) && (!b || aligned16(b)) && (!c || aligned16(c));
}

extern "C" int pmu_bn_train_fwd_f32(const float* y, const float* gamma, const float* beta, float eps, int relu,
                                    float momentum, float* run_mean, float* run_var, float* mean, float* var,
                                    float* a, double* ws, int B, int C, int64_t HW, void* stream) {
  PMU_CHECK_ARG(y && gamma && beta && mean && var && a && ws, "pmu_bn_train_fwd_f32: null pointer");
  PMU_CHECK_ARG(B > 0 && C > 0 && C <= 65535 && HW > 0 && (int64_t)B * C < (1ll << 31), "pmu_bn_train_fwd_f32: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t chunk; int nch;
  stat_chunks(C, HW, &chunk, &nch);
  const bool vec = rows_vec(HW, y, a);
  if (vec) bn_stats_kernel<true><<<dim3(nch, C), 256, 0, st>>>(y, B, C, HW, chunk, ws);
  else bn_stats_kernel<false><<<dim3(nch, C), 256, 0, st>>>(y, B, C, HW, chunk, ws);
  PMU_LAUNCH_CHECK();
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws, C, (double)B * (double)HW, mean, var, run_mean, run_var, momentum);
  PMU_LAUNCH_CHECK();
  const unsigned gx = (unsigned)std::min<int64_t>(cdiv64(HW, vec ? 1024 : 256), 64);
  if (vec) bn_act_kernel<true><<<dim3(B * C, gx), 256, 0, st>>>(y, mean, var, gamma, beta, eps, relu, a, C, HW);
  else bn_act_kernel<false><<<dim3(B * C, gx), 256, 0, st>>>(y, mean, var, gamma, beta, eps, relu, a, C, HW);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_bn_train_bwd_bias_f32(const float* da, const float* y, const float* mean, const float* var,
                                         const float* gamma, const float* beta, float eps, int relu, float* dy,
                                         float* dgamma, float* dbeta, float* dbias, double* ws, int B, int C,
                                         int64_t HW, void* stream) {
  PMU_CHECK_ARG(da && y && mean && var && gamma && beta && dy && ws, "pmu_bn_train_bwd_f32: null pointer");
  PMU_CHECK_ARG(B > 0 && C > 0 && C <= 65535 && HW > 0 && (int64_t)B * C < (1ll << 31), "pmu_bn_train_bwd_f32: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (dbias ? 3 : 2) * C, st));
  int64_t chunk; int nch;
  stat_chunks(C, HW, &chunk, &nch);
  const bool vec = rows_vec(HW, da, y, dy);
  if (vec) bn_bwd_reduce_kernel<true><<<dim3(nch, C), 256, 0, st>>>(da, y, mean, var, gamma, beta, eps, relu, B, C, HW, chunk, ws);
  else bn_bwd_reduce_kernel<false><<<dim3(nch, C), 256, 0, st>>>(da, y, mean, var, gamma, beta, eps, relu, B, C, HW, chunk, ws);
  PMU_LAUNCH_CHECK();
  const unsigned gx = (unsigned)std::min<int64_t>(cdiv64(HW, vec ? 1024 : 256), 64);
  double* bias_acc = dbias ? ws + 2 * (size_t)C : nullptr;
  if (vec)
    bn_bwd_apply_kernel<true><<<dim3(B * C, gx), 256, 0, st>>>(da, y, mean, var, gamma, beta, eps, relu, ws,
                                                              (double)B * (double)HW, dy, dgamma, dbeta, bias_acc, C, HW);
  else
    bn_bwd_apply_kernel<false><<<dim3(B * C, gx), 256, 0, st>>>(da, y, mean, var, gamma, beta, eps, relu, ws,
                                                               (double)B * (double)HW, dy, dgamma, dbeta, bias_acc, C, HW);
  PMU_LAUNCH_CHECK();
  if (dbias) {
    channel_sum_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(bias_acc, 1, C, dbias);
    PMU_LAUNCH_CHECK();
  }
  return PMU_OK;
}

extern "C" int pmu_bn_train_bwd_f32(const float* da, const float* y, const float* mean, const float* var,
                                    const float* gamma, const float* beta, float eps, int relu, float* dy,
                                    float* dgamma, float* dbeta, double* ws, int B, int C, int64_t HW,
                                    void* stream) {
  return pmu_bn_train_bwd_bias_f32(da, y, mean, var, gamma, beta, eps, relu, dy, dgamma, dbeta, nullptr, ws, B, C, HW, stream);
}

extern "C" int pmu_channel_sums_f32(const float* x, float* out, double* ws, int B, int C, int64_t HW, void* stream) {
  PMU_CHECK_ARG(x && out && ws && B > 0 && C > 0 && HW > 0, "pmu_channel_sums_f32: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t chunk; int nch;
  stat_chunks(C, HW, &chunk, &nch);
  if (rows_vec(HW, x)) bn_stats_kernel<true><<<dim3(nch, C), 256, 0, st>>>(x, B, C, HW, chunk, ws);
  else bn_stats_kernel<false><<<dim3(nch, C), 256, 0, st>>>(x, B, C, HW, chunk, ws);
  PMU_LAUNCH_CHECK();
  channel_sum_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws, 2, C, out);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_row_sums_f32(const float* x, float* out, int64_t rows, int64_t n, void* stream) {
  PMU_CHECK_ARG(x && out && rows > 0 && rows < (1ll << 31) && n > 0, "pmu_row_sums_f32: bad argument");
  row_sum_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

namespace pmu {
// conv3x3 weight gradient for the 1- and 2-channel input layers (inc.c1 of the U-Net, the first layers of the prior /
// posterior encoders): the 32 x 32 (co x ci) tile of the general kernel would waste 15/16 of its ci lanes there.
// grid (chunks, Cout): a block owns one output channel and a range of pixel quads over all images; a thread walks quads
// (4 consecutive pixels of a row: one 128-bit dy load, 3 x 6 x-taps per input channel) with 9 * CIN running sums in
// registers; block reduction, then one fp32 atomicAdd per weight (same accumulation contract as the general kernel).
template <int CIN>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_smallcin_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                              const float* __restrict__ dy, float* __restrict__ dw, int H, int W, int Cout,
                              int64_t quads, int64_t quads_per_block) {
  const int co = blockIdx.y;
  const int W4 = W >> 2;
  const int64_t HW = (int64_t)H * W;
  const int64_t lo = (int64_t)blockIdx.x * quads_per_block;
  const int64_t hi = (lo + quads_per_block < quads) ? lo + quads_per_block : quads;
  float acc[CIN][3][3];
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[c][k / 3][k % 3] = 0.f;
  for (int64_t q = lo + threadIdx.x; q < hi; q += 256) {
    const int w0 = (int)(q % W4) * 4;
    const int64_t r = q / W4;
    const int h = (int)(r % H);
    const int64_t b = r / H;
    const float4 d4 = *reinterpret_cast<const float4*>(dy + (b * Cout + co) * HW + (int64_t)h * W + w0);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float* xs = (c < C0) ? x0 + (b * C0 + c) * HW : x1 + (b * C1 + (c - C0)) * HW;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hh = h + ky - 1;
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int ww = w0 + j - 1;
          v[j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xs + (int64_t)hh * W + ww) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[c][ky][kx] = fmaf(d[j], v[j + kx], acc[c][ky][kx]);
      }
    }
  }
  __shared__ float red[8][CIN * 9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float v = warp_sum(acc[c][k / 3][k % 3]);
      if (lane == 0) red[warp][c * 9 + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < CIN * 9) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    atomicAdd(dw + (int64_t)co * CIN * 9 + threadIdx.x, v);       // dw[co][ci][ky][kx], ci * 9 + ky * 3 + kx
  }
}
}  // namespace pmu

extern "C" int pmu_conv3x3_wgrad_f32(const float* x0, int C0, const float* x1, int C1, const float* dy, float* dw,
                                     int B, int H, int W, int Cout, void* stream) {
  PMU_CHECK_ARG(x0 && dy && dw && (C1 == 0 || x1), "pmu_conv3x3_wgrad_f32: null pointer");
  PMU_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cout > 0 && C0 > 0 && C1 >= 0, "pmu_conv3x3_wgrad_f32: bad shape");
  const int Cin = C0 + C1;
  if (Cin <= 2 && W % 4 == 0 && aligned16(dy) && Cout <= 65535) {
    const int64_t quads = (int64_t)B * H * (W / 4);
    const int64_t want = std::max<int64_t>(1, cdiv64(8ll * sm_count(), Cout));
    const int64_t qpb = std::max<int64_t>(256, cdiv64(quads, want));
    const dim3 grid((unsigned)cdiv64(quads, qpb), Cout);
    if (Cin == 1)
      conv3x3_wgrad_smallcin_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, dy, dw, H, W, Cout, quads, qpb);
    else
      conv3x3_wgrad_smallcin_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, dy, dw, H, W, Cout, quads, qpb);
    PMU_LAUNCH_CHECK();
    return PMU_OK;
  }
  const int tiles_x = cdiv(W, WG_TW), tiles_y = cdiv(H, WG_TH);
  const int gx = cdiv(Cout, WG_CO) * cdiv(Cin, WG_CI);
  const int units = B * tiles_x * tiles_y;
  int gy = std::max(1, std::min(units, cdiv(6 * sm_count(), gx)));
  const int upb = cdiv(units, gy);
  gy = cdiv(units, upb);
  conv3x3_wgrad_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, dy, dw, B, H, W, Cout, tiles_x,
                                                                      tiles_y, upb);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

namespace pmu {
// conv1x1 wgrad; dw has row stride ldw (layer 0 of fcomb writes into the [F, F+L] weight gradient)
__global__ void __launch_bounds__(256)
conv1x1_wgrad_ld_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int ldw,
                        int B, int Cin, int Cout, int64_t HW, int chunks_per_img, int units_per_block) {
  __shared__ float x_s[32 * (W1_P + 1)];
  __shared__ float d_s[32 * (W1_P + 1)];
  const int ci_tiles = (Cin + 31) / 32;
  const int co0 = (blockIdx.x / ci_tiles) * 32, ci0 = (blockIdx.x % ci_tiles) * 32;
  const int tid = threadIdx.x, tci = tid & 15, tco = tid >> 4;
  const int units = B * chunks_per_img;
  const int u_lo = blockIdx.y * units_per_block, u_hi = min(units, u_lo + units_per_block);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int u = u_lo; u < u_hi; ++u) {
    const int b = u / chunks_per_img;
    const int64_t p0 = (int64_t)(u % chunks_per_img) * W1_P;
    __syncthreads();
    for (int idx = tid; idx < 32 * W1_P; idx += 256) {
      const int c = idx / W1_P, p = idx % W1_P;
      float xv = 0.f, dv = 0.f;
      if (p0 + p < HW) {
        if (ci0 + c < Cin) xv = __ldg(x + ((int64_t)b * Cin + ci0 + c) * HW + p0 + p);
        if (co0 + c < Cout) dv = __ldg(dy + ((int64_t)b * Cout + co0 + c) * HW + p0 + p);
      }
      x_s[c * (W1_P + 1) + p] = xv;
      d_s[c * (W1_P + 1) + p] = dv;
    }
    __syncthreads();
    const float* xa = x_s + (2 * tci) * (W1_P + 1);
    const float* da_ = d_s + (2 * tco) * (W1_P + 1);
#pragma unroll 8
    for (int p = 0; p < W1_P; ++p) {
      const float a0 = xa[p], a1 = xa[W1_P + 1 + p], d0 = da_[p], d1 = da_[W1_P + 1 + p];
      acc[0][0] = fmaf(d0, a0, acc[0][0]); acc[0][1] = fmaf(d0, a1, acc[0][1]);
      acc[1][0] = fmaf(d1, a0, acc[1][0]); acc[1][1] = fmaf(d1, a1, acc[1][1]);
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int co = co0 + 2 * tco + a, ci = ci0 + 2 * tci + b;
      if (co < Cout && ci < Cin) atomicAdd(dw + (int64_t)co * ldw + ci, acc[a][b]);
    }
}
}  // namespace pmu

extern "C" int pmu_conv1x1_wgrad_f32(const float* x, const float* dy, float* dw, int ldw, int B, int Cin, int Cout,
                                     int64_t HW, void* stream) {
  PMU_CHECK_ARG(x && dy && dw && B > 0 && Cin > 0 && Cout > 0 && HW > 0 && ldw >= Cin, "pmu_conv1x1_wgrad_f32: bad argument");
  const int gx = cdiv(Cout, 32) * cdiv(Cin, 32);
  const int cpi = (int)cdiv64(HW, W1_P);
  const int units = B * cpi;
  int gy = std::max(1, std::min(units, cdiv(6 * sm_count(), gx)));
  const int upb = cdiv(units, gy);
  gy = cdiv(units, upb);
  conv1x1_wgrad_ld_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(x, dy, dw, ldw, B, Cin, Cout, HW, cpi, upb);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pool2_bwd_f32(const float* x, const float* dy, float* dx, int B, int C, int H, int W, int mode,
                                 void* stream) {
  PMU_CHECK_ARG(dy && dx && (mode == PMU_POOL_AVG_CEIL || x), "pmu_pool2_bwd_f32: null pointer");
  PMU_CHECK_ARG(mode == PMU_POOL_MAX || mode == PMU_POOL_AVG_CEIL, "pmu_pool2_bwd_f32: unknown mode %d", mode);
  PMU_CHECK_SUPPORTED(mode == PMU_POOL_AVG_CEIL || (H % 2 == 0 && W % 2 == 0),
                      "pmu_pool2_bwd_f32: max-pool backward needs even H, W (got %dx%d)", H, W);
  const int Ho = (mode == PMU_POOL_MAX) ? H / 2 : (H + 1) / 2, Wo = (mode == PMU_POOL_MAX) ? W / 2 : (W + 1) / 2;
  const int64_t total = (int64_t)B * C * Ho * Wo;
  pool2_bwd_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, H, W, Ho, Wo, mode, total);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_convt2x2_dgrad_f32(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H,
                                      int W, void* stream) {
  PMU_CHECK_ARG(dy && w && dx && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "pmu_convt2x2_dgrad_f32: bad argument");
  PMU_CHECK_ARG(aligned16(w) && (reinterpret_cast<uintptr_t>(dy) & 7u) == 0, "pmu_convt2x2_dgrad_f32: alignment");
  const int64_t HW = (int64_t)H * W;
  convt2x2_dgrad_kernel<<<dim3((unsigned)cdiv64(HW, 128), cdiv(Cin, 8), B), 128, 0, (cudaStream_t)stream>>>(dy, w, dx, Cin,
                                                                                                       Cout, H, W);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_convt2x2_wgrad_f32(const float* x, const float* dy, float* dw, int B, int Cin, int Cout, int H,
                                      int W, void* stream) {
  PMU_CHECK_ARG(x && dy && dw && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "pmu_convt2x2_wgrad_f32: bad argument");
  const int gx = cdiv(Cin, 32) * cdiv(Cout, 16);
  const int cpi = (int)cdiv64((int64_t)H * W, CT_P);
  const int units = B * cpi;
  int gy = std::max(1, std::min(units, cdiv(6 * sm_count(), gx)));
  const int upb = cdiv(units, gy);
  gy = cdiv(units, upb);
  convt2x2_wgrad_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(x, dy, dw, B, Cin, Cout, H, W, cpi, upb);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_relu_bwd_f32(const float* a, const float* dy, float* dx, int64_t n, void* stream) {
  PMU_CHECK_ARG(a && dy && dx && n > 0, "pmu_relu_bwd_f32: bad argument");
  relu_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, dy, dx, n);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_add_f32(float* dst, const float* src, int64_t n, void* stream) {
  PMU_CHECK_ARG(dst && src && n > 0, "pmu_add_f32: bad argument");
  add_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_ce_bwd_f32(const float* logits, const float* target, float scale, float* dlogits, int B, int C,
                              int64_t HW, void* stream) {
  PMU_CHECK_ARG(logits && target && dlogits && B > 0 && B <= 65535 && C > 0 && HW > 0, "pmu_ce_bwd_f32: bad argument");
  ce_bwd_kernel<<<dim3((unsigned)cdiv64(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(logits, target, scale, dlogits, C, HW);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_kl_bwd_f32(const float* mu_q, const float* ls_q, const float* mu_p, const float* ls_p, float scale,
                              float* dmu_q, float* dls_q, float* dmu_p, float* dls_p, int B, int L, void* stream) {
  PMU_CHECK_ARG(mu_q && ls_q && mu_p && ls_p && dmu_q && dls_q && dmu_p && dls_p && B > 0 && L > 0, "pmu_kl_bwd_f32: bad argument");
  kl_bwd_kernel<<<cdiv(B * L, 128), 128, 0, (cudaStream_t)stream>>>(mu_q, ls_q, mu_p, ls_p, scale, dmu_q, dls_q, dmu_p,
                                                                    dls_p, B * L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_gauss_head_bwd_f32(const float* enc, const float* w, const float* dmu, const float* dls, float* denc,
                                      float* dw, float* db, int B, int C, int h, int w_, int L, void* stream) {
  PMU_CHECK_ARG(enc && w && dmu && dls && denc && dw && db && B > 0 && C > 0 && h > 0 && w_ > 0 && L > 0 && 2 * L <= 256,
                "pmu_gauss_head_bwd_f32: bad argument");
  gauss_head_bwd_kernel<<<B, 256, 2 * L * sizeof(float), (cudaStream_t)stream>>>(enc, w, dmu, dls, denc, dw, db, C, h * w_, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fcomb_zbias_f32(const float* z, const float* w0, const float* b0, float* zb, int B, int F, int L,
                                   void* stream) {
  PMU_CHECK_ARG(z && w0 && b0 && zb && B > 0 && F > 0 && L > 0, "pmu_fcomb_zbias_f32: bad argument");
  fcomb_zbias_kernel<<<cdiv(B * F, 128), 128, 0, (cudaStream_t)stream>>>(z, w0, b0, zb, B, F, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fcomb_zbias_bwd_f32(const float* rs, const float* z, const float* w0, float* dz, float* dw0,
                                       float* db0, int B, int F, int L, void* stream) {
  PMU_CHECK_ARG(rs && z && w0 && dz && dw0 && db0 && B > 0 && F > 0 && L > 0, "pmu_fcomb_zbias_bwd_f32: bad argument");
  fcomb_zbias_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rs, z, w0, dz, dw0, db0, B, F, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_conv1x1_bb_f32(const float* x, const float* w, int ldw, const float* bias, int bias_bstride, float* y,
                                  int B, int Cin, int Cout, int64_t HW, int relu, void* stream) {
  PMU_CHECK_ARG(x && w && y && B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && HW > 0 && ldw >= Cin, "pmu_conv1x1_bb_f32: bad argument");
  PMU_CHECK_SUPPORTED(Cin <= 1536, "pmu_conv1x1_bb_f32: Cin <= 1536 (the 8 x Cin weight slab is staged in 48 KB of shared memory)");
  const dim3 grid((unsigned)cdiv64(HW, 1024), cdiv(Cout, 8), B);
  const size_t smem = (size_t)Cin * 8 * sizeof(float);
  if (HW % 4 == 0 && aligned16(x) && aligned16(y))
    conv1x1_bb_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, ldw, bias, bias_bstride, y, Cin, Cout, HW, relu);
  else
    conv1x1_bb_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, ldw, bias, bias_bstride, y, Cin, Cout, HW, relu);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
