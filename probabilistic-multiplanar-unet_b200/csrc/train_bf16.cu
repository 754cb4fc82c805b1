// Training step, tensor-core mode, elementwise / reduction side on bf16 NHWC activations — the layout the tcgen05
// convolutions (conv_tc.cu forward / data gradient, wgrad_tc.cu weight gradient) read and write, so that no layout or
// precision cast sits between two GEMMs of the step (round 1: fp32 NCHW BatchNorm / pooling kernels bracketed every
// GEMM with nchw_f32 <-> nhwc_bf16 casts: 4.2 of the 20 ms of kernel time, and the BatchNorm passes moved 12-28 B per
// element where these move 6-10).
//
// Replaces, in train() mode (train.py:94-110): nn.BatchNorm2d + ReLU forward / backward (unet_parts.py:16-20,
// probabilistic_unet.py:39-45), nn.MaxPool2d(2) / nn.AvgPool2d(2, 2, ceil_mode=True) backward (unet_parts.py:33,
// probabilistic_unet.py:36), the skip-connection gradient add of torch.cat (unet_parts.py:65), the bias gradient of
// nn.ConvTranspose2d (unet_parts.py:52) and the Gaussian head's backward (probabilistic_unet.py:97-108).
//
// All tensors are [npix = B*H*W][C] bf16 with C % 8 == 0 (a thread moves 8 channels = 16 B).  Per-channel reductions:
// a block walks a contiguous range of pixels, thread = (8-channel group, pixel lane); fp32 partials per thread, combined
// through shared memory, then ONE fp64 atomic per channel and block (order-independent to fp32).
//
// Per-channel reductions (BatchNorm statistics, its backward sums, bias gradients) are two-level and deterministic:
// a block walks a contiguous range of pixels, thread = (8-channel group, pixel lane), fp32 partials per thread ->
// per-block partials through shared memory -> ONE row of a [blocks][NV][C] fp32 workspace; the finalize kernel adds
// the rows in fp64.  (Round 2a used fp64 atomics from ~1200 blocks onto the same 2*C addresses: the statistics pass of
// a 67 MB tensor ran at 1.6 TB/s, bound by the atomics of one or two L2 slices, and needed a memset in front.)
// Loads are issued four at a time before they are consumed (one 16-byte load in flight per thread does not cover
// the HBM latency at 2048 threads per SM).
//
// Why the BatchNorm statistics are a separate pass and not in the convolution epilogue: the epilogue owns a
// [128 pixel][64 channel] tile with one pixel row per thread, so per-channel sums are a reduction ACROSS its 128
// threads — 62 shuffles + 62 adds per 32 columns and thread, or 32 shared-memory loads + 128 fp32 ops per thread from
// the staging tile — which doubles an epilogue that already is the critical path of the 64-channel layers; this pass
// reads the stored tensor (still in L2 for most layers) once.
#include "pmu_common.cuh"
#include "h16.cuh"

namespace pmu {

constexpr int TB_THREADS = 256;
constexpr int RED_ROW = TB_THREADS + 4;          // padded row of the block-reduction scratch: conflict-free both ways
constexpr int RED_MAX_BLOCKS = PMU_RED_MAX_BLOCKS;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = unpack16<false>(u[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack16_rn<false>(f[0], f[1]), pack16_rn<false>(f[2], f[3]), pack16_rn<false>(f[4], f[5]), pack16_rn<false>(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// per-block channel partials: acc[v * 8 + k] of thread (g, pl) -> out[v * C + g * 8 + k] = sum over the block's pixel lanes
template <int NV>
__device__ __forceinline__ void block_channel_partials(const float (&acc)[8 * NV], float* part, int G, int planes, int C,
                                                       float* __restrict__ out) {
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8 * NV; ++j) part[j * RED_ROW + t] = acc[j];
  __syncthreads();
  for (int c = t; c < C; c += TB_THREADS) {
    const int g = c >> 3, k = c & 7;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float* row = part + (v * 8 + k) * RED_ROW + g;
      float s = 0.f;
      for (int pl = 0; pl < planes; ++pl) s += row[pl * G];
      out[v * C + c] = s;
    }
  }
}

// Second level: out[v][c] (fp64) = sum over the nb rows of ws[row][v][c].  Block = 32 channels x 32 row lanes (1024 threads),
// four rows in flight per thread: the pass is a chain of dependent L2 round trips, so its time is rows / (lanes x depth)
// latencies (the first version, 8 row lanes and one load in flight, took 15-18 us per BatchNorm layer: 1.15 ms per step).
constexpr int FIN_THREADS = 1024, FIN_LANES = 32;
template <int NV>
__device__ __forceinline__ void sum_partials(const float* __restrict__ ws, int nb, int C, int c, int r, double (&s)[NV]) {
#pragma unroll
  for (int v = 0; v < NV; ++v) s[v] = 0.0;
  if (c >= C) return;
  int b = r;
  for (; b + 3 * FIN_LANES < nb; b += 4 * FIN_LANES) {
    float t[4][NV];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v) t[u][v] = __ldg(ws + ((int64_t)(b + u * FIN_LANES) * NV + v) * C + c);
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] += ((double)t[0][v] + (double)t[1][v]) + ((double)t[2][v] + (double)t[3][v]);
  }
  for (; b < nb; b += FIN_LANES)
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] += (double)__ldg(ws + ((int64_t)b * NV + v) * C + c);
}
template <int NV>
__device__ __forceinline__ bool reduce_rows(double (&s)[NV], double* sh /* [FIN_LANES][32][NV] */) {
  const int cl = threadIdx.x & 31, r = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) sh[(r * 32 + cl) * NV + v] = s[v];
  __syncthreads();
  if (r != 0) return false;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double a = 0.0;
    for (int q = 0; q < FIN_LANES; ++q) a += sh[(q * 32 + cl) * NV + v];
    s[v] = a;
  }
  return true;
}

// ws[block] = {sum y, sum y^2} per channel over the block's pixels
__global__ void __launch_bounds__(TB_THREADS)
bn_stats_nhwc_kernel(const uint4* __restrict__ y, int64_t npix, int C, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  auto eat = [&](const uint4& v) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] += f[k]; a[8 + k] = fmaf(f[k], f[k], a[8 + k]); }
  };
  int64_t p = lo + pl;
  for (; p + 3 * planes < hi; p += 4 * planes) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream_u4(y + (p + (int64_t)u * planes) * G + g);
#pragma unroll
    for (int u = 0; u < 4; ++u) eat(v[u]);
  }
  for (; p < hi; p += planes) eat(ldg_stream_u4(y + p * G + g));
  block_channel_partials<2>(a, part, G, planes, C, ws + (int64_t)blockIdx.x * 2 * C);
}

__global__ void __launch_bounds__(FIN_THREADS)
bn_finalize_nhwc_kernel(const float* __restrict__ ws, int nb, int C, double n, float eps, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ var,
                        float* __restrict__ run_mean, float* __restrict__ run_var, float momentum,
                        float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ double sh[FIN_LANES * 32 * 2];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s[2];
  sum_partials<2>(ws, nb, C, c, threadIdx.x >> 5, s);
  if (!reduce_rows<2>(s, sh) || c >= C) return;
  const double m = s[0] / n;
  double v = s[1] / n - m * m;
  if (v < 0) v = 0;
  mean[c] = (float)m;
  var[c] = (float)v;
  if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
  if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(n > 1 ? v * n / (n - 1) : v);
  const float inv = 1.f / sqrtf((float)v + eps);
  scale[c] = gamma[c] * inv;                     // a = y * scale + shift
  shift[c] = beta[c] - (float)m * gamma[c] * inv;
}

// BatchNorm finalize from the per-channel fp64 {sum, sum^2} the convolution epilogue accumulated (pmu_conv_gemm_bnstats_bf16)
__global__ void bn_finalize_acc_nhwc_kernel(const double* __restrict__ acc, int C, double n, float eps, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ var,
                                            float* __restrict__ run_mean, float* __restrict__ run_var, float momentum,
                                            float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[2 * c] / n;
  double v = acc[2 * c + 1] / n - m * m;
  if (v < 0) v = 0;
  mean[c] = (float)m;
  var[c] = (float)v;
  if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
  if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(n > 1 ? v * n / (n - 1) : v);
  const float inv = 1.f / sqrtf((float)v + eps);
  scale[c] = gamma[c] * inv;
  shift[c] = beta[c] - (float)m * gamma[c] * inv;
}

// a = [relu](y * scale + shift), elementwise over [npix][C].  The grid stride (gridDim * 256) is a multiple of G = C / 8,
// so a thread keeps ONE channel group for its whole loop: scale / shift live in registers.
__global__ void __launch_bounds__(TB_THREADS)
bn_act_nhwc_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                   uint4* __restrict__ a, int64_t total, int G) {
  const int64_t i0 = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * TB_THREADS;
  const int g = (int)(i0 % G);
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + g * 8 + k); sh[k] = __ldg(shift + g * 8 + k); }
  auto act = [&](const uint4& v) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float z = fmaf(f[k], sc[k], sh[k]); f[k] = relu ? fmaxf(z, 0.f) : z; }
    return pack8(f);
  };
  int64_t i = i0;
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream_u4(y + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) a[i + u * stride] = act(v[u]);
  }
  for (; i < total; i += stride) a[i] = act(ldg_stream_u4(y + i));
}

// backward pass 1: dz = da * (z > 0), z = y * scale + shift;  ws[block] = {sum dz, sum dz * xhat}, xhat = (y - mean) * inv
__global__ void __launch_bounds__(TB_THREADS)
bn_bwd_reduce_nhwc_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y, const float* __restrict__ mean,
                          const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float eps, int relu, int64_t npix, int C, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  float m[8], inv[8], sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    m[k] = __ldg(mean + c); inv[k] = 1.f / sqrtf(__ldg(var + c) + eps);
    sc[k] = __ldg(gamma + c) * inv[k]; sh[k] = __ldg(beta + c) - m[k] * sc[k];      // the forward's z = y * scale + shift
  }
  auto eat = [&](const uint4& vy, const uint4& vd) {
    float fy[8], fd[8];
    unpack8(vy, fy);
    unpack8(vd, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d = fd[k];
      if (relu && !(fmaf(fy[k], sc[k], sh[k]) > 0.f)) d = 0.f;
      a[k] += d; a[8 + k] = fmaf(d, (fy[k] - m[k]) * inv[k], a[8 + k]);
    }
  };
  int64_t p = lo + pl;
  for (; p + planes < hi; p += 2 * planes) {
    const uint4 y0 = ldg_stream_u4(y + p * G + g), d0 = ldg_stream_u4(da + p * G + g);
    const uint4 y1 = ldg_stream_u4(y + (p + planes) * G + g), d1 = ldg_stream_u4(da + (p + planes) * G + g);
    eat(y0, d0);
    eat(y1, d1);
  }
  for (; p < hi; p += planes) eat(ldg_stream_u4(y + p * G + g), ldg_stream_u4(da + p * G + g));
  block_channel_partials<2>(a, part, G, planes, C, ws + (int64_t)blockIdx.x * 2 * C);
}

// backward finalize: dbeta = sum dz, dgamma = sum dz * xhat, and the per-channel coefficients of the elementwise pass:
//   dy = gamma * inv * (dz - mean(dz) - xhat * mean(dz * xhat)) = A * dz + Bc * y + Cc
//   coef[0] = scale (= A), coef[1] = shift (ReLU mask: y * scale + shift > 0), coef[2] = Bc, coef[3] = Cc
__global__ void __launch_bounds__(FIN_THREADS)
bn_bwd_finalize_nhwc_kernel(const float* __restrict__ ws, int nb, int C, double n, float eps, const float* __restrict__ mean,
                            const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  __shared__ double sh[FIN_LANES * 32 * 2];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s[2];
  sum_partials<2>(ws, nb, C, c, threadIdx.x >> 5, s);
  if (!reduce_rows<2>(s, sh) || c >= C) return;
  if (dbeta) dbeta[c] = (float)s[0];
  if (dgamma) dgamma[c] = (float)s[1];
  const float m = mean[c], inv = 1.f / sqrtf(var[c] + eps), scale = gamma[c] * inv;
  const float m1 = (float)(s[0] / n), m2 = (float)(s[1] / n);
  const float bc = -scale * inv * m2;
  coef[c] = scale;
  coef[C + c] = beta[c] - m * scale;
  coef[2 * C + c] = bc;
  coef[3 * C + c] = -scale * m1 - bc * m;
}

// backward pass 2, elementwise: dy = A * dz + Bc * y + Cc
__global__ void __launch_bounds__(TB_THREADS)
bn_bwd_apply_nhwc_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y, const float* __restrict__ coef, int relu,
                         uint4* __restrict__ dy, int64_t total, int C) {
  const int G = C >> 3;
  const int64_t i0 = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * TB_THREADS;
  const int g = (int)(i0 % G);                   // constant over the loop: the grid stride is a multiple of G
  float sc[8], sh[8], bc[8], cc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    sc[k] = __ldg(coef + c); sh[k] = __ldg(coef + C + c); bc[k] = __ldg(coef + 2 * C + c); cc[k] = __ldg(coef + 3 * C + c);
  }
  auto apply = [&](const uint4& vy, const uint4& vd) {
    float fy[8], fd[8];
    unpack8(vy, fy);
    unpack8(vd, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d = fd[k];
      if (relu && !(fmaf(fy[k], sc[k], sh[k]) > 0.f)) d = 0.f;
      fd[k] = fmaf(sc[k], d, fmaf(bc[k], fy[k], cc[k]));
    }
    return pack8(fd);
  };
  int64_t i = i0;
  for (; i + stride < total; i += 2 * stride) {
    const uint4 y0 = ldg_stream_u4(y + i), d0 = ldg_stream_u4(da + i);
    const uint4 y1 = ldg_stream_u4(y + i + stride), d1 = ldg_stream_u4(da + i + stride);
    dy[i] = apply(y0, d0);
    dy[i + stride] = apply(y1, d1);
  }
  for (; i < total; i += stride) dy[i] = apply(ldg_stream_u4(y + i), ldg_stream_u4(da + i));
}

// per-channel sums of a bf16 [nseg][npix][C] tensor, one result row per segment (bias gradients: the transposed
// convolution's and the fcomb layers' with nseg = 1; the per-slice latent bias of fcomb layer 0 with nseg = B)
__global__ void __launch_bounds__(TB_THREADS)
channel_sums_nhwc_kernel(const uint4* __restrict__ x, int64_t npix, int C, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  x += (int64_t)blockIdx.y * npix * G;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  auto eat = [&](const uint4& v) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += f[k];
  };
  int64_t p = lo + pl;
  for (; p + 3 * planes < hi; p += 4 * planes) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream_u4(x + (p + (int64_t)u * planes) * G + g);
#pragma unroll
    for (int u = 0; u < 4; ++u) eat(v[u]);
  }
  for (; p < hi; p += planes) eat(ldg_stream_u4(x + p * G + g));
  block_channel_partials<1>(a, part, G, planes, C, ws + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * C);
}
__global__ void __launch_bounds__(FIN_THREADS)
channel_sums_finalize_nhwc_kernel(const float* __restrict__ ws, int nb, int C, float* __restrict__ out) {
  __shared__ double sh[FIN_LANES * 32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s[1];
  sum_partials<1>(ws + (int64_t)blockIdx.y * nb * C, nb, C, c, threadIdx.x >> 5, s);
  if (!reduce_rows<1>(s, sh) || c >= C) return;
  out[(int64_t)blockIdx.y * C + c] = (float)s[0];
}

// 2x2 pooling backward, NHWC bf16: thread = (output window, 8 channels).  MAX: the gradient goes to the first maximum of
// the window in row-major order (torch; the fp32 kernel's rule); AVG_CEIL: dy / (number of in-bounds taps).
__global__ void __launch_bounds__(TB_THREADS)
pool2_bwd_nhwc_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W, int Ho,
                      int Wo, int G, int mode, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * TB_THREADS) {
    const int g = (int)(i % G);
    int64_t r = i / G;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int64_t b = r / Ho;
    const int y0 = 2 * oy, x0 = 2 * ox;
    float gd[8];
    unpack8(__ldg(dy + i), gd);
    const int ny = (y0 + 1 < H) ? 2 : 1, nx = (x0 + 1 < W) ? 2 : 1;
    auto at = [&](int a, int c) { return ((b * H + y0 + a) * W + x0 + c) * G + g; };
    if (mode == PMU_POOL_MAX) {
      float v[4][8];
      for (int t = 0; t < 4; ++t) unpack8(__ldg(x + at(t >> 1, t & 1)), v[t]);       // MaxPool2d(2): even H, W
      float o[4][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int best = 0;
        float bv = v[0][k];
#pragma unroll
        for (int t = 1; t < 4; ++t) if (v[t][k] > bv) { bv = v[t][k]; best = t; }
#pragma unroll
        for (int t = 0; t < 4; ++t) o[t][k] = (t == best) ? gd[k] : 0.f;
      }
      for (int t = 0; t < 4; ++t) dx[at(t >> 1, t & 1)] = pack8(o[t]);
    } else {
      const float inv = 1.f / (float)(ny * nx);
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = gd[k] * inv;
      const uint4 pk = pack8(o);
      for (int a = 0; a < ny; ++a)
        for (int c = 0; c < nx; ++c) dx[at(a, c)] = pk;
    }
  }
}

__global__ void __launch_bounds__(TB_THREADS) add_bf16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x; i < n8; i += (int64_t)gridDim.x * TB_THREADS) {
    float a[8], b[8];
    unpack8(dst[i], a);
    unpack8(__ldg(src + i), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    dst[i] = pack8(a);
  }
}

// Gaussian head backward on a bf16 NHWC encoder map (probabilistic_unet.py:97-108: mean over H, W, then 1x1 conv to 2L):
//   denc[b,p,c] = (sum_o w[o][c] * d[b][o]) / hw   (the same for every pixel p);  dw[o][c] += d[b][o] * mean_p enc[b,p,c];
//   db[o] += d[b][o].  d = (dmu | dlog_sigma).  Block = (slice b, 64 channels): 8 channel groups x 32 pixel lanes
//   (round 2a: one block per slice, a thread per channel walking the pixels with 2-byte accesses — 136 us on 8 blocks).
__global__ void __launch_bounds__(TB_THREADS)
gauss_head_bwd_nhwc_kernel(const uint4* __restrict__ enc, const float* __restrict__ w, const float* __restrict__ dmu,
                           const float* __restrict__ dls, uint4* __restrict__ denc, float* __restrict__ dw,
                           float* __restrict__ db, int C, int hw, int L) {
  __shared__ float part[8 * RED_ROW];
  __shared__ uint32_t dq[32];
  const int b = blockIdx.x, G = C >> 3;
  const int g = threadIdx.x & 7, pl = threadIdx.x >> 3, gg = blockIdx.y * 8 + g;     // global channel group
  const bool live = gg < G;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (live)
    for (int p = pl; p < hw; p += 32) {
      float f[8];
      unpack8(__ldg(enc + ((int64_t)b * hw + p) * G + gg), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[j * RED_ROW + threadIdx.x] = a[j];
  if (threadIdx.x < 32) dq[threadIdx.x] = 0u;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cg = threadIdx.x >> 3, k = threadIdx.x & 7, c = blockIdx.y * 64 + threadIdx.x;
    float s = 0.f;
    for (int q = 0; q < 32; ++q) s += part[k * RED_ROW + q * 8 + cg];
    const float m = s / (float)hw;
    if (c < C) {
      float dv = 0.f;
      for (int o = 0; o < 2 * L; ++o) {
        const float d = (o < L) ? __ldg(dmu + b * L + o) : __ldg(dls + b * L + o - L);
        dv = fmaf(__ldg(w + (int64_t)o * C + c), d, dv);
        atomicAdd(dw + (int64_t)o * C + c, d * m);
      }
      reinterpret_cast<__nv_bfloat16*>(dq)[threadIdx.x] = __float2bfloat16(dv / (float)hw);
    }
  }
  __syncthreads();
  if (live) {
    const uint4 q = make_uint4(dq[g * 4], dq[g * 4 + 1], dq[g * 4 + 2], dq[g * 4 + 3]);
    for (int p = pl; p < hw; p += 32) denc[((int64_t)b * hw + p) * G + gg] = q;
  }
  if (blockIdx.y == 0 && threadIdx.x < 2 * L) {
    const int o = threadIdx.x;
    atomicAdd(db + o, (o < L) ? __ldg(dmu + b * L + o) : __ldg(dls + b * L + o - L));
  }
}

// ---- weight layouts of the tensor-core training step ------------------------------------------------------------------
// The module keeps nn.Conv2d weights as fp32 OIHW [Cout][Cin][3][3]; the tcgen05 GEMMs want bf16 K-major operands:
// forward  wf[co][tap][ci], data gradient (= convolution with the transposed, flipped weights) wd[ci][8 - tap][co].
// Round 2a built both with torch (to(bf16) / permute / flip / contiguous: ~230 small kernels and 2 ms per step for 69 M
// parameters); here one kernel reads a 32 x 32 (co, ci) tile once and writes both layouts with 4-byte stores.
__device__ __forceinline__ void pack_conv3x3_tile(const float* __restrict__ w, uint32_t* __restrict__ wf, uint32_t* __restrict__ wd,
                                                  int Cout, int Cin, int co0, int ci0, float (*s)[289]) {
  // a row of the tile = 32 input channels x 9 taps = 288 contiguous floats (16-byte aligned: ci0 % 32 == 0): 72 float4
#pragma unroll 3
  for (int i = threadIdx.x; i < 32 * 72; i += TB_THREADS) {
    const int r = i / 72, c = (i - r * 72) * 4;
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(w + ((int64_t)(co0 + r) * Cin + ci0) * 9 + c));
    s[r][c] = v.x; s[r][c + 1] = v.y; s[r][c + 2] = v.z; s[r][c + 3] = v.w;
  }
  __syncthreads();
  if (wf)
    for (int i = threadIdx.x; i < 32 * 9 * 16; i += TB_THREADS) {      // (co_l, tap, ci pair), ci fastest
      const int cp = i & 15, t = (i >> 4) % 9, r = i / 144;
      wf[(((int64_t)(co0 + r) * 9 + t) * Cin + ci0) / 2 + cp] = pack16_rn<false>(s[r][(2 * cp) * 9 + t], s[r][(2 * cp + 1) * 9 + t]);
    }
  if (wd)
    for (int i = threadIdx.x; i < 32 * 9 * 16; i += TB_THREADS) {      // (ci_l, tap, co pair), co fastest
      const int cp = i & 15, t = (i >> 4) % 9, c = i / 144;
      wd[(((int64_t)(ci0 + c) * 9 + (8 - t)) * Cout + co0) / 2 + cp] = pack16_rn<false>(s[2 * cp][c * 9 + t], s[2 * cp + 1][c * 9 + t]);
    }
}
__global__ void __launch_bounds__(TB_THREADS)
pack_conv3x3_kernel(const float* __restrict__ w, uint32_t* __restrict__ wf, uint32_t* __restrict__ wd, int Cout, int Cin) {
  __shared__ float s[32][289];
  pack_conv3x3_tile(w, wf, wd, Cout, Cin, blockIdx.y * 32, blockIdx.x * 32, s);
}
// every 3x3 layer of a network in ONE launch: table[l] = {w, wf, wd, Cout, Cin, first tile} (int64 each), a block = one
// 32 x 32 tile of one layer (35 launches of 5-30 us, most of them latency, become one pass at memory rate)
__global__ void __launch_bounds__(TB_THREADS)
pack_conv3x3_multi_kernel(const int64_t* __restrict__ table, int nlayers) {
  __shared__ float s[32][289];
  int l = 0;
  while (l + 1 < nlayers && (int64_t)blockIdx.x >= __ldg(table + (l + 1) * 6 + 5)) ++l;
  const int64_t* e = table + l * 6;
  const int Cout = (int)__ldg(e + 3), Cin = (int)__ldg(e + 4);
  const int t = (int)((int64_t)blockIdx.x - __ldg(e + 5)), tiles_ci = Cin / 32;
  pack_conv3x3_tile(reinterpret_cast<const float*>(__ldg(e)), reinterpret_cast<uint32_t*>(__ldg(e + 1)),
                    reinterpret_cast<uint32_t*>(__ldg(e + 2)), Cout, Cin, (t / tiles_ci) * 32, (t % tiles_ci) * 32, s);
}
// weight gradient of the tcgen05 wgrad kernel, fp32 [Cout][9][Cin] -> the parameter's OIHW [Cout][Cin][3][3]
__global__ void __launch_bounds__(TB_THREADS)
unpack_wgrad3x3_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cin) {
  __shared__ float s[9][65];
  const int co = blockIdx.y, ci0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 9 * 64; i += TB_THREADS) {
    const int t = i >> 6, c = i & 63;
    s[t][c] = __ldg(dwp + ((int64_t)co * 9 + t) * Cin + ci0 + c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 64; i += TB_THREADS) {
    const int c = i / 9, t = i - c * 9;
    dw[((int64_t)co * Cin + ci0) * 9 + i] = s[t][c];
  }
}

// ---- Fcomb head of the training step on bf16 NHWC (probabilistic_unet.py:137-146, 167-181) ---------------------------
// Its three 64 -> 64 1x1 layers run on the tcgen05 GEMMs (forward, data and weight gradients); these kernels are the
// ends of the chain: the last layer F -> n_classes (fp32 NCHW logits for the cross entropy), its backward, and the ReLU
// mask between two data-gradient GEMMs.  (Round 2a ran the whole head on fp32 NCHW CUDA-core kernels: 3.3 of 16.7 ms.)

// logits[b][k][p] = sum_c w[k][c] h[b,p,c] + bias[k].  8 threads per pixel (one 16-byte load each), 3 xor-shuffles per class.
template <int MAXC>
__global__ void __launch_bounds__(TB_THREADS)
fcomb_last_fwd_kernel(const uint4* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ logits, int64_t npix, int64_t HW, int F, int C) {
  const int G = F >> 3;                                  // 8 threads cover 64 channels; F > 64: loop over groups of 8 threads' reach
  const int g = threadIdx.x & 7;
  const int64_t p = ((int64_t)blockIdx.x * TB_THREADS + threadIdx.x) >> 3;
  float acc[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) acc[k] = 0.f;
  if (p < npix)
    for (int gg = g; gg < G; gg += 8) {
      float f[8];
      unpack8(ldg_stream_u4(h + p * G + gg), f);
#pragma unroll
      for (int k = 0; k < MAXC; ++k)
        if (k < C) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (int64_t)k * F + gg * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (int64_t)k * F + gg * 8 + 4));
          acc[k] = fmaf(f[0], w0.x, fmaf(f[1], w0.y, fmaf(f[2], w0.z, fmaf(f[3], w0.w, acc[k]))));
          acc[k] = fmaf(f[4], w1.x, fmaf(f[5], w1.y, fmaf(f[6], w1.z, fmaf(f[7], w1.w, acc[k]))));
        }
    }
#pragma unroll
  for (int k = 0; k < MAXC; ++k) {
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 4);
  }
  if (p < npix) {
    const int64_t b = p / HW, q = p - b * HW;
#pragma unroll
    for (int k = 0; k < MAXC; ++k)
      if (k < C && (k & 7) == g) logits[(b * C + k) * HW + q] = acc[k] + (bias ? __ldg(bias + k) : 0.f);
  }
}

// backward of that layer: dh[b,p,c] = (h > 0) * sum_k w[k][c] dl[b][k][p] (bf16 NHWC, the ReLU of the layer below folded in);
// ws[block] = per-block partials of dw[k][c] = sum_p dl[k] * h[c]  (finalised by channel_sums_finalize with C rows of F)
template <int MAXC>
__global__ void __launch_bounds__(TB_THREADS)
fcomb_last_bwd_kernel(const uint4* __restrict__ h, const float* __restrict__ dl, const float* __restrict__ w,
                      uint4* __restrict__ dh, int64_t npix, int64_t HW, int F, int C, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float part[];
  const int G = F >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float wr[MAXC][8], a[8 * MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) { wr[k][j] = (k < C) ? __ldg(w + (int64_t)k * F + g * 8 + j) : 0.f; a[k * 8 + j] = 0.f; }
  for (int64_t p = lo + pl; p < hi; p += planes) {
    const int64_t b = p / HW, q = p - b * HW;
    float f[8], d[MAXC], o[8];
    unpack8(ldg_stream_u4(h + p * G + g), f);
#pragma unroll
    for (int k = 0; k < MAXC; ++k) d[k] = (k < C) ? __ldg(dl + (b * C + k) * HW + q) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < MAXC; ++k) { v = fmaf(wr[k][j], d[k], v); a[k * 8 + j] = fmaf(d[k], f[j], a[k * 8 + j]); }
      o[j] = (f[j] > 0.f) ? v : 0.f;
    }
    dh[p * G + g] = pack8(o);
  }
  // per-block partials, one row of [MAXC][F] per block (rows k >= C are zero)
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8 * MAXC; ++j) part[j * RED_ROW + t] = a[j];
  __syncthreads();
  float* out = ws + (int64_t)blockIdx.x * C * F;
  for (int i = t; i < C * F; i += TB_THREADS) {
    const int k = i / F, c = i - k * F, cg = c >> 3, j = c & 7;
    const float* row = part + (k * 8 + j) * RED_ROW + cg;
    float sum = 0.f;
    for (int q = 0; q < planes; ++q) sum += row[q * G];
    out[i] = sum;
  }
}

// d = (h > 0) ? d : 0, in place: the ReLU between two data-gradient GEMMs of the Fcomb chain
__global__ void __launch_bounds__(TB_THREADS) relu_mask_bf16_kernel(uint4* __restrict__ d, const uint4* __restrict__ h, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x; i < n8; i += (int64_t)gridDim.x * TB_THREADS) {
    float a[8], b[8];
    unpack8(d[i], a);
    unpack8(ldg_stream_u4(h + i), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (b[k] > 0.f) ? a[k] : 0.f;
    d[i] = pack8(a);
  }
}

// ---- weight gradient of the three first layers (Cin = 1: U-Net, prior; Cin = 2: posterior) ---------------------------
// dw[co][ci][ky][kx] = sum_{b,h,w} dy[b,h,w,co] * x_ci[b,h+ky-1,w+kx-1]: dy is the bf16 NHWC gradient of the step, x the
// fp32 NCHW image (x0) / mask (x1).  Same two-level reduction as above with nine values per channel; blockIdx.y = ci.
// (Round 2a cast dy to fp32 NCHW and ran the fp32 small-Cin kernel: 0.54 ms per step for 1728 output values.)
__global__ void __launch_bounds__(TB_THREADS)
wgrad_smallcin_nhwc_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const uint4* __restrict__ dy, int64_t npix,
                           int H, int W, int C, int64_t chunk, float* __restrict__ ws) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  const float* __restrict__ x = blockIdx.y == 0 ? x0 : x1;
  const int64_t HW = (int64_t)H * W;
  float a[72];
#pragma unroll
  for (int j = 0; j < 72; ++j) a[j] = 0.f;
  for (int64_t p = lo + pl; p < hi; p += planes) {
    const int64_t b = p / HW;
    const int r = (int)(p - b * HW), h = r / W, w = r - h * W;
    float f[8], t[9];
    unpack8(ldg_stream_u4(dy + p * G + g), f);
    const float* xb = x + b * HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int hh = h + ky - 1, ww = w + kx - 1;
        t[ky * 3 + kx] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + (int64_t)hh * W + ww) : 0.f;
      }
#pragma unroll
    for (int v = 0; v < 9; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[v * 8 + k] = fmaf(f[k], t[v], a[v * 8 + k]);
  }
  block_channel_partials<9>(a, part, G, planes, C, ws + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 9 * C);
}
__global__ void __launch_bounds__(FIN_THREADS)
wgrad_smallcin_finalize_kernel(const float* __restrict__ ws, int nb, int C, int Cin, float* __restrict__ dw) {
  __shared__ double sh[FIN_LANES * 32];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);            // i = v * C + c over the 9 * C sums of input channel blockIdx.y
  double s[1];
  sum_partials<1>(ws + (int64_t)blockIdx.y * nb * 9 * C, nb, 9 * C, i, threadIdx.x >> 5, s);
  if (!reduce_rows<1>(s, sh) || i >= 9 * C) return;
  const int v = i / C, c = i - v * C;
  dw[((int64_t)c * Cin + blockIdx.y) * 9 + v] = (float)s[0];
}

// blocks of a per-channel reduction over npix pixels: at least 32 pixels per block (the per-block partial rows must stay
// small next to the tensor), at most PMU_RED_MAX_BLOCKS (4 per SM)
static inline void reduce_grid(int64_t npix, int64_t* chunk, unsigned* blocks) {
  const int64_t nb = std::max<int64_t>(1, std::min<int64_t>(RED_MAX_BLOCKS, npix / 32));
  *chunk = cdiv64(npix, nb);
  *blocks = (unsigned)cdiv64(npix, *chunk);
}
// grid of an elementwise pass over `total` 16-byte vectors: ~8 vectors per thread
static inline unsigned ew_blocks(int64_t total) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv64(total, (int64_t)TB_THREADS * 8), (int64_t)sm_count() * 8));
}

}  // namespace pmu

using namespace pmu;

#define PMU_NHWC_ARGS(name)                                                                                            \
  PMU_CHECK_ARG(npix > 0 && C > 0, name ": bad shape");                                                                \
  PMU_CHECK_SUPPORTED(C % 8 == 0 && C <= 2048 && (C / 8) <= TB_THREADS && TB_THREADS % (C / 8) == 0,                   \
                      name ": C must be a multiple of 8 with C/8 dividing 256 (got %d)", C)

extern "C" int pmu_bn_train_fwd_nhwc_bf16(const void* y, const float* gamma, const float* beta, float eps, int relu,
                                          float momentum, float* run_mean, float* run_var, float* mean, float* var,
                                          void* a, float* ws, float* scale_shift, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(y && gamma && beta && mean && var && a && ws && scale_shift, "pmu_bn_train_fwd_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_bn_train_fwd_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(y) && aligned16(a) && aligned16(scale_shift), "pmu_bn_train_fwd_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, &chunk, &blocks);
  bn_stats_nhwc_kernel<<<blocks, TB_THREADS, 16 * RED_ROW * sizeof(float), st>>>(reinterpret_cast<const uint4*>(y), npix, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  bn_finalize_nhwc_kernel<<<cdiv(C, 32), FIN_THREADS, 0, st>>>(ws, (int)blocks, C, (double)npix, eps, gamma, beta, mean, var, run_mean,
                                                             run_var, momentum, scale_shift, scale_shift + C);
  PMU_LAUNCH_CHECK();
  const int64_t total = npix * (C / 8);
  bn_act_nhwc_kernel<<<ew_blocks(total), TB_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(y), scale_shift, scale_shift + C, relu,
                                                             reinterpret_cast<uint4*>(a), total, C / 8);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_bn_train_bwd_nhwc_bf16(const void* da, const void* y, const float* mean, const float* var, const float* gamma,
                                          const float* beta, float eps, int relu, void* dy, float* dgamma, float* dbeta,
                                          float* ws, float* coef, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(da && y && mean && var && gamma && beta && dy && ws && coef, "pmu_bn_train_bwd_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_bn_train_bwd_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(y) && aligned16(da) && aligned16(dy), "pmu_bn_train_bwd_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, &chunk, &blocks);
  bn_bwd_reduce_nhwc_kernel<<<blocks, TB_THREADS, 16 * RED_ROW * sizeof(float), st>>>(
      reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(y), mean, var, gamma, beta, eps, relu, npix, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  bn_bwd_finalize_nhwc_kernel<<<cdiv(C, 32), FIN_THREADS, 0, st>>>(ws, (int)blocks, C, (double)npix, eps, mean, var, gamma, beta, dgamma,
                                                                 dbeta, coef);
  PMU_LAUNCH_CHECK();
  const int64_t total = npix * (C / 8);
  bn_bwd_apply_nhwc_kernel<<<ew_blocks(total), TB_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(y),
                                                                   coef, relu, reinterpret_cast<uint4*>(dy), total, C);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_channel_sums_nhwc_bf16(const void* x, float* out, float* ws, int nseg, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(x && out && ws && nseg > 0 && nseg <= 65535, "pmu_channel_sums_nhwc_bf16: bad arguments");
  PMU_NHWC_ARGS("pmu_channel_sums_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(x), "pmu_channel_sums_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, &chunk, &blocks);
  if (nseg > 1) {      // the segments share the PMU_RED_MAX_BLOCKS rows of the workspace
    const int64_t nb = std::max<int64_t>(1, std::min<int64_t>(blocks, RED_MAX_BLOCKS / nseg));
    chunk = cdiv64(npix, nb);
    blocks = (unsigned)cdiv64(npix, chunk);
  }
  PMU_CHECK_SUPPORTED((int64_t)blocks * nseg <= RED_MAX_BLOCKS, "pmu_channel_sums_nhwc_bf16: nseg = %d exceeds the workspace rows", nseg);
  channel_sums_nhwc_kernel<<<dim3(blocks, nseg), TB_THREADS, 8 * RED_ROW * sizeof(float), st>>>(reinterpret_cast<const uint4*>(x), npix, C,
                                                                                              chunk, ws);
  PMU_LAUNCH_CHECK();
  channel_sums_finalize_nhwc_kernel<<<dim3(cdiv(C, 32), nseg), FIN_THREADS, 0, st>>>(ws, (int)blocks, C, out);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pool2_bwd_nhwc_bf16(const void* x, const void* dy, void* dx, int B, int H, int W, int C, int mode, void* stream) {
  PMU_CHECK_ARG(dy && dx && (mode == PMU_POOL_AVG_CEIL || x), "pmu_pool2_bwd_nhwc_bf16: null pointer");
  PMU_CHECK_ARG(mode == PMU_POOL_MAX || mode == PMU_POOL_AVG_CEIL, "pmu_pool2_bwd_nhwc_bf16: unknown mode %d", mode);
  PMU_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "pmu_pool2_bwd_nhwc_bf16: bad shape");
  PMU_CHECK_SUPPORTED(C % 8 == 0, "pmu_pool2_bwd_nhwc_bf16: C must be a multiple of 8 (got %d)", C);
  PMU_CHECK_SUPPORTED(mode == PMU_POOL_AVG_CEIL || (H % 2 == 0 && W % 2 == 0), "pmu_pool2_bwd_nhwc_bf16: MaxPool2d(2) backward needs even H, W");
  const int Ho = (mode == PMU_POOL_MAX) ? H / 2 : (H + 1) / 2, Wo = (mode == PMU_POOL_MAX) ? W / 2 : (W + 1) / 2;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  const unsigned g = (unsigned)std::min<int64_t>(cdiv64(total, TB_THREADS), (int64_t)sm_count() * 16);
  pool2_bwd_nhwc_kernel<<<g, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy),
                                                                   reinterpret_cast<uint4*>(dx), H, W, Ho, Wo, C / 8, mode, total);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_add_bf16(void* dst, const void* src, int64_t n, void* stream) {
  PMU_CHECK_ARG(dst && src && n >= 0, "pmu_add_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(n % 8 == 0 && aligned16(dst) && aligned16(src), "pmu_add_bf16: n must be a multiple of 8, pointers 16-byte aligned");
  if (n == 0) return PMU_OK;
  const unsigned g = (unsigned)std::min<int64_t>(cdiv64(n / 8, TB_THREADS), (int64_t)sm_count() * 16);
  add_bf16_kernel<<<g, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n / 8);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_gauss_head_bwd_nhwc_bf16(const void* enc, const float* w, const float* dmu, const float* dls, void* denc,
                                            float* dw, float* db, int B, int C, int h, int w_, int L, void* stream) {
  PMU_CHECK_ARG(enc && w && dmu && dls && denc && dw && db, "pmu_gauss_head_bwd_nhwc_bf16: null pointer");
  PMU_CHECK_ARG(B > 0 && C > 0 && h > 0 && w_ > 0 && L > 0 && 2 * L <= TB_THREADS, "pmu_gauss_head_bwd_nhwc_bf16: bad shape");
  PMU_CHECK_SUPPORTED(C % 8 == 0 && aligned16(enc) && aligned16(denc),
                      "pmu_gauss_head_bwd_nhwc_bf16: C must be a multiple of 8 (got %d), tensors 16-byte aligned", C);
  gauss_head_bwd_nhwc_kernel<<<dim3(B, cdiv(C, 64)), TB_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(enc), w, dmu, dls, reinterpret_cast<uint4*>(denc), dw, db, C, h * w_, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pack_conv3x3_weights_bf16(const float* w, void* wf, void* wd, int Cout, int Cin, void* stream) {
  PMU_CHECK_ARG(w && (wf || wd) && Cout > 0 && Cin > 0, "pmu_pack_conv3x3_weights_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(Cout % 32 == 0 && Cin % 32 == 0 && Cout / 32 <= 65535,
                      "pmu_pack_conv3x3_weights_bf16: channel counts must be multiples of 32 (got %d, %d)", Cout, Cin);
  pack_conv3x3_kernel<<<dim3(Cin / 32, Cout / 32), TB_THREADS, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<uint32_t*>(wf), reinterpret_cast<uint32_t*>(wd), Cout, Cin);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_unpack_conv3x3_wgrad_f32(const float* dwp, float* dw, int Cout, int Cin, void* stream) {
  PMU_CHECK_ARG(dwp && dw && Cout > 0 && Cin > 0, "pmu_unpack_conv3x3_wgrad_f32: bad arguments");
  PMU_CHECK_SUPPORTED(Cin % 64 == 0 && Cout <= 65535, "pmu_unpack_conv3x3_wgrad_f32: Cin must be a multiple of 64 (got %d)", Cin);
  unpack_wgrad3x3_kernel<<<dim3(Cin / 64, Cout), TB_THREADS, 0, (cudaStream_t)stream>>>(dwp, dw, Cin);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fcomb_last_fwd_bf16(const void* h, const float* w, const float* bias, float* logits, int B, int64_t HW, int F,
                                       int C, void* stream) {
  PMU_CHECK_ARG(h && w && logits && B > 0 && HW > 0 && F > 0 && C > 0, "pmu_fcomb_last_fwd_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(F % 8 == 0 && C <= 8 && aligned16(h) && aligned16(w), "pmu_fcomb_last_fwd_bf16: F %% 8 == 0, n_classes <= 8 (got %d, %d)", F, C);
  const int64_t npix = (int64_t)B * HW;
  const unsigned grid = (unsigned)cdiv64(npix * 8, TB_THREADS);
  if (C <= 4) fcomb_last_fwd_kernel<4><<<grid, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(h), w, bias, logits, npix, HW, F, C);
  else fcomb_last_fwd_kernel<8><<<grid, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(h), w, bias, logits, npix, HW, F, C);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fcomb_last_bwd_bf16(const void* h, const float* dlogits, const float* w, void* dh, float* dw, float* ws, int B,
                                       int64_t HW, int F, int C, void* stream) {
  PMU_CHECK_ARG(h && dlogits && w && dh && dw && ws && B > 0 && HW > 0, "pmu_fcomb_last_bwd_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(F % 8 == 0 && (F / 8) <= TB_THREADS && TB_THREADS % (F / 8) == 0 && C > 0 && C <= 4 && aligned16(h) && aligned16(dh),
                      "pmu_fcomb_last_bwd_bf16: F/8 must divide 256, n_classes <= 4 (got %d, %d)", F, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t npix = (int64_t)B * HW;
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, &chunk, &blocks);
  fcomb_last_bwd_kernel<4><<<blocks, TB_THREADS, 32 * RED_ROW * sizeof(float), st>>>(
      reinterpret_cast<const uint4*>(h), dlogits, w, reinterpret_cast<uint4*>(dh), npix, HW, F, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  channel_sums_finalize_nhwc_kernel<<<dim3(cdiv(C * F, 32), 1), FIN_THREADS, 0, st>>>(ws, (int)blocks, C * F, dw);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_relu_mask_bf16(void* d, const void* h, int64_t n, void* stream) {
  PMU_CHECK_ARG(d && h && n >= 0, "pmu_relu_mask_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(n % 8 == 0 && aligned16(d) && aligned16(h), "pmu_relu_mask_bf16: n must be a multiple of 8, pointers 16-byte aligned");
  if (n == 0) return PMU_OK;
  const unsigned g = (unsigned)std::min<int64_t>(cdiv64(n / 8, TB_THREADS), (int64_t)sm_count() * 16);
  relu_mask_bf16_kernel<<<g, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(d), reinterpret_cast<const uint4*>(h), n / 8);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_conv3x3_wgrad_smallcin_bf16(const float* x0, const float* x1, const void* dy, float* dw, float* ws, int B, int H,
                                               int W, int Cout, void* stream) {
  PMU_CHECK_ARG(x0 && dy && dw && ws && B > 0 && H > 0 && W > 0 && Cout > 0, "pmu_conv3x3_wgrad_smallcin_bf16: bad arguments");
  const int C = Cout;
  const int64_t npix = (int64_t)B * H * W;
  PMU_NHWC_ARGS("pmu_conv3x3_wgrad_smallcin_bf16");
  PMU_CHECK_ARG(aligned16(dy), "pmu_conv3x3_wgrad_smallcin_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cin = x1 ? 2 : 1;
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, &chunk, &blocks);
  const int64_t nb = std::max<int64_t>(1, std::min<int64_t>(blocks, RED_MAX_BLOCKS / 2));      // rows: Cin * nb <= PMU_RED_MAX_BLOCKS
  chunk = cdiv64(npix, nb);
  blocks = (unsigned)cdiv64(npix, chunk);
  const int dyn = 72 * RED_ROW * (int)sizeof(float);
  PMU_CUDA(cudaFuncSetAttribute(wgrad_smallcin_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  wgrad_smallcin_nhwc_kernel<<<dim3(blocks, Cin), TB_THREADS, dyn, st>>>(x0, x1, reinterpret_cast<const uint4*>(dy), npix, H, W, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  wgrad_smallcin_finalize_kernel<<<dim3(cdiv(9 * C, 32), Cin), FIN_THREADS, 0, st>>>(ws, (int)blocks, C, Cin, dw);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_bn_train_fwd_stats_nhwc_bf16(const void* y, const double* stats, const float* gamma, const float* beta, float eps,
                                                int relu, float momentum, float* run_mean, float* run_var, float* mean, float* var,
                                                void* a, float* scale_shift, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(y && stats && gamma && beta && mean && var && a && scale_shift, "pmu_bn_train_fwd_stats_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_bn_train_fwd_stats_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(y) && aligned16(a), "pmu_bn_train_fwd_stats_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  bn_finalize_acc_nhwc_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats, C, (double)npix, eps, gamma, beta, mean, var, run_mean, run_var, momentum,
                                                           scale_shift, scale_shift + C);
  PMU_LAUNCH_CHECK();
  const int64_t total = npix * (C / 8);
  bn_act_nhwc_kernel<<<ew_blocks(total), TB_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(y), scale_shift, scale_shift + C, relu,
                                                             reinterpret_cast<uint4*>(a), total, C / 8);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pack_conv3x3_weights_multi_bf16(const int64_t* table, int nlayers, int64_t total_tiles, void* stream) {
  PMU_CHECK_ARG(table && nlayers > 0 && total_tiles > 0 && total_tiles < (1ll << 31), "pmu_pack_conv3x3_weights_multi_bf16: bad arguments");
  pack_conv3x3_multi_kernel<<<(unsigned)total_tiles, TB_THREADS, 0, (cudaStream_t)stream>>>(table, nlayers);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
