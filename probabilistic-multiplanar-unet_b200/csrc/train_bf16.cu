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
// Why the BatchNorm statistics are a separate pass and not in the convolution epilogue: the epilogue owns a
// [128 pixel][64 channel] tile with one pixel row per thread, so per-channel sums are a reduction ACROSS its 128
// threads — 62 shuffles + 62 adds per 32 columns and thread, or 32 shared-memory loads + 128 fp32 ops per thread from
// the staging tile — which doubles an epilogue that already is the critical path of the 64-channel layers; this pass
// reads the stored tensor once at HBM rate (~0.25 ms for all 38 BatchNorm layers of a batch-8 step).
#include "pmu_common.cuh"
#include "h16.cuh"

namespace pmu {

constexpr int TB_THREADS = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = unpack16<false>(u[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack16_rn<false>(f[0], f[1]), pack16_rn<false>(f[2], f[3]), pack16_rn<false>(f[4], f[5]), pack16_rn<false>(f[6], f[7]));
}

// block-level per-channel reduction of NV values per channel: part[TB_THREADS][8 * NV] in shared memory
template <int NV>
__device__ __forceinline__ void channel_reduce(const float (&acc)[8 * NV], float* part, int G, int planes, int C,
                                               double* __restrict__ out, int stride) {
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8 * NV; ++j) part[t * (8 * NV) + j] = acc[j];
  __syncthreads();
  for (int c = t; c < C; c += TB_THREADS) {
    const int g = c >> 3, k = c & 7;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float s = 0.f;
      for (int pl = 0; pl < planes; ++pl) s += part[(pl * G + g) * (8 * NV) + v * 8 + k];
      atomicAdd(out + (int64_t)c * stride + v, (double)s);
    }
  }
}

// acc[c][0] += sum y, acc[c][1] += sum y^2
__global__ void __launch_bounds__(TB_THREADS)
bn_stats_nhwc_kernel(const uint4* __restrict__ y, int64_t npix, int C, int64_t chunk, double* __restrict__ acc) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  if (pl < planes) {
#pragma unroll 4
    for (int64_t p = lo + pl; p < hi; p += planes) {
      float f[8];
      unpack8(__ldg(y + p * G + g), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] += f[k]; a[8 + k] = fmaf(f[k], f[k], a[8 + k]); }
    }
  }
  channel_reduce<2>(a, part, G, planes, C, acc, 2);
}

__global__ void bn_finalize_nhwc_kernel(const double* __restrict__ acc, int C, double n, float eps, const float* __restrict__ gamma,
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
  scale[c] = gamma[c] * inv;                     // a = y * scale + shift
  shift[c] = beta[c] - (float)m * gamma[c] * inv;
}

// a = [relu](y * scale + shift), elementwise over [npix][C].  The grid stride (gridDim * 256) is a multiple of G = C / 8,
// so a thread keeps ONE channel group for its whole loop: scale / shift live in registers.
__global__ void __launch_bounds__(TB_THREADS)
bn_act_nhwc_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                   uint4* __restrict__ a, int64_t total, int G) {
  const int64_t i0 = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x;
  const int g = (int)(i0 % G);
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + g * 8 + k); sh[k] = __ldg(shift + g * 8 + k); }
  for (int64_t i = i0; i < total; i += (int64_t)gridDim.x * TB_THREADS) {
    float f[8];
    unpack8(__ldg(y + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float v = fmaf(f[k], sc[k], sh[k]); f[k] = relu ? fmaxf(v, 0.f) : v; }
    a[i] = pack8(f);
  }
}

// backward pass 1: dz = da * (z > 0), z = y * scale + shift;  acc[c] += {sum dz, sum dz * xhat}, xhat = (y - mean) * inv
__global__ void __launch_bounds__(TB_THREADS)
bn_bwd_reduce_nhwc_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y, const float* __restrict__ mean,
                          const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float eps, int relu, int64_t npix, int C, int64_t chunk, double* __restrict__ acc) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  if (pl < planes) {
    float m[8], inv[8], gm[8], bt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      m[k] = __ldg(mean + c); inv[k] = 1.f / sqrtf(__ldg(var + c) + eps); gm[k] = __ldg(gamma + c); bt[k] = __ldg(beta + c);
    }
#pragma unroll 2
    for (int64_t p = lo + pl; p < hi; p += planes) {
      float fy[8], fd[8];
      unpack8(__ldg(y + p * G + g), fy);
      unpack8(__ldg(da + p * G + g), fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (fy[k] - m[k]) * inv[k];
        float d = fd[k];
        if (relu && !(fmaf(fy[k], gm[k] * inv[k], bt[k] - m[k] * gm[k] * inv[k]) > 0.f)) d = 0.f;   // the forward's z = y * scale + shift
        a[k] += d; a[8 + k] = fmaf(d, xh, a[8 + k]);
      }
    }
  }
  channel_reduce<2>(a, part, G, planes, C, acc, 2);
}

// backward pass 2: dy = gamma * inv * (dz - mean(dz) - xhat * mean(dz * xhat)), elementwise; block 0 also writes
// dgamma = sum dz * xhat, dbeta = sum dz
__global__ void __launch_bounds__(TB_THREADS)
bn_bwd_apply_nhwc_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y, const float* __restrict__ mean,
                         const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, int relu, const double* __restrict__ acc, double n, uint4* __restrict__ dy,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t total, int C) {
  const int G = C >> 3;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += TB_THREADS) {
      if (dbeta) dbeta[c] = (float)acc[2 * c];
      if (dgamma) dgamma[c] = (float)acc[2 * c + 1];
    }
  const int64_t i0 = (int64_t)blockIdx.x * TB_THREADS + threadIdx.x;
  const int g = (int)(i0 % G);                   // constant over the loop: the grid stride is a multiple of G
  float m[8], inv[8], gm[8], bt[8], m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    m[k] = __ldg(mean + c); inv[k] = 1.f / sqrtf(__ldg(var + c) + eps); gm[k] = __ldg(gamma + c); bt[k] = __ldg(beta + c);
    m1[k] = (float)(acc[2 * c] / n); m2[k] = (float)(acc[2 * c + 1] / n);
  }
  for (int64_t i = i0; i < total; i += (int64_t)gridDim.x * TB_THREADS) {
    float fy[8], fd[8];
    unpack8(__ldg(y + i), fy);
    unpack8(__ldg(da + i), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (fy[k] - m[k]) * inv[k];
      float d = fd[k];
      if (relu && !(fmaf(fy[k], gm[k] * inv[k], bt[k] - m[k] * gm[k] * inv[k]) > 0.f)) d = 0.f;   // the forward's z = y * scale + shift
      fd[k] = gm[k] * inv[k] * (d - m1[k] - xh * m2[k]);
    }
    dy[i] = pack8(fd);
  }
}

// per-channel sums of a bf16 NHWC tensor (bias gradient of the transposed convolution): acc[c] += sum x
__global__ void __launch_bounds__(TB_THREADS)
channel_sums_nhwc_kernel(const uint4* __restrict__ x, int64_t npix, int C, int64_t chunk, double* __restrict__ acc) {
  extern __shared__ float part[];
  const int G = C >> 3, planes = TB_THREADS / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < npix) ? lo + chunk : npix;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (pl < planes)
    for (int64_t p = lo + pl; p < hi; p += planes) {
      float f[8];
      unpack8(__ldg(x + p * G + g), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
  channel_reduce<1>(a, part, G, planes, C, acc, 1);
}
__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)a[i];
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
//   db[o] += d[b][o].  d = (dmu | dlog_sigma).  Block per slice b, thread per channel.
__global__ void __launch_bounds__(TB_THREADS)
gauss_head_bwd_nhwc_kernel(const __nv_bfloat16* __restrict__ enc, const float* __restrict__ w, const float* __restrict__ dmu,
                           const float* __restrict__ dls, __nv_bfloat16* __restrict__ denc, float* __restrict__ dw,
                           float* __restrict__ db, int C, int hw, int L) {
  const int b = blockIdx.x;
  const float inv = 1.f / (float)hw;
  for (int c = threadIdx.x; c < C; c += TB_THREADS) {
    float m = 0.f;
    for (int p = 0; p < hw; ++p) m += __bfloat162float(enc[((int64_t)b * hw + p) * C + c]);
    m *= inv;
    float dv = 0.f;
    for (int o = 0; o < 2 * L; ++o) {
      const float d = (o < L) ? __ldg(dmu + b * L + o) : __ldg(dls + b * L + o - L);
      dv = fmaf(__ldg(w + (int64_t)o * C + c), d, dv);
      atomicAdd(dw + (int64_t)o * C + c, d * m);
    }
    const __nv_bfloat16 q = __float2bfloat16(dv * inv);
    for (int p = 0; p < hw; ++p) denc[((int64_t)b * hw + p) * C + c] = q;
  }
  if (threadIdx.x < 2 * L) {
    const int o = threadIdx.x;
    atomicAdd(db + o, (o < L) ? __ldg(dmu + b * L + o) : __ldg(dls + b * L + o - L));
  }
}

static inline void reduce_grid(int64_t npix, int C, int64_t* chunk, unsigned* blocks) {
  // enough blocks for the machine, chunks of at least 64 pixels per plane-sweep
  int64_t nb = std::min<int64_t>((int64_t)sm_count() * 8, std::max<int64_t>(1, npix * (int64_t)C / (8 * 256 * 8)));
  *chunk = cdiv64(npix, nb);
  *blocks = (unsigned)cdiv64(npix, *chunk);
}

}  // namespace pmu

using namespace pmu;

#define PMU_NHWC_ARGS(name)                                                                                            \
  PMU_CHECK_ARG(npix > 0 && C > 0, name ": bad shape");                                                                \
  PMU_CHECK_SUPPORTED(C % 8 == 0 && C <= 2048 && (C / 8) <= TB_THREADS && TB_THREADS % (C / 8) == 0,                   \
                      name ": C must be a multiple of 8 with C/8 dividing 256 (got %d)", C)

extern "C" int pmu_bn_train_fwd_nhwc_bf16(const void* y, const float* gamma, const float* beta, float eps, int relu,
                                          float momentum, float* run_mean, float* run_var, float* mean, float* var,
                                          void* a, double* ws, float* scale_shift, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(y && gamma && beta && mean && var && a && ws && scale_shift, "pmu_bn_train_fwd_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_bn_train_fwd_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(y) && aligned16(a) && aligned16(scale_shift), "pmu_bn_train_fwd_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, C, &chunk, &blocks);
  bn_stats_nhwc_kernel<<<blocks, TB_THREADS, TB_THREADS * 16 * sizeof(float), st>>>(reinterpret_cast<const uint4*>(y), npix, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  bn_finalize_nhwc_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws, C, (double)npix, eps, gamma, beta, mean, var, run_mean, run_var, momentum,
                                                       scale_shift, scale_shift + C);
  PMU_LAUNCH_CHECK();
  const int64_t total = npix * (C / 8);
  const unsigned g = (unsigned)std::min<int64_t>(cdiv64(total, TB_THREADS), (int64_t)sm_count() * 16);
  bn_act_nhwc_kernel<<<g, TB_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(y), scale_shift, scale_shift + C, relu,
                                              reinterpret_cast<uint4*>(a), total, C / 8);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_bn_train_bwd_nhwc_bf16(const void* da, const void* y, const float* mean, const float* var, const float* gamma,
                                          const float* beta, float eps, int relu, void* dy, float* dgamma, float* dbeta,
                                          double* ws, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(da && y && mean && var && gamma && beta && dy && ws, "pmu_bn_train_bwd_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_bn_train_bwd_nhwc_bf16");
  PMU_CHECK_ARG(aligned16(y) && aligned16(da) && aligned16(dy), "pmu_bn_train_bwd_nhwc_bf16: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, C, &chunk, &blocks);
  bn_bwd_reduce_nhwc_kernel<<<blocks, TB_THREADS, TB_THREADS * 16 * sizeof(float), st>>>(
      reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(y), mean, var, gamma, beta, eps, relu, npix, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  const int64_t total = npix * (C / 8);
  const unsigned g = (unsigned)std::min<int64_t>(cdiv64(total, TB_THREADS), (int64_t)sm_count() * 16);
  bn_bwd_apply_nhwc_kernel<<<g, TB_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(y), mean, var,
                                                    gamma, beta, eps, relu, ws, (double)npix, reinterpret_cast<uint4*>(dy), dgamma,
                                                    dbeta, total, C);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_channel_sums_nhwc_bf16(const void* x, float* out, double* ws, int64_t npix, int C, void* stream) {
  PMU_CHECK_ARG(x && out && ws, "pmu_channel_sums_nhwc_bf16: null pointer");
  PMU_NHWC_ARGS("pmu_channel_sums_nhwc_bf16");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * C, st));
  int64_t chunk; unsigned blocks;
  reduce_grid(npix, C, &chunk, &blocks);
  channel_sums_nhwc_kernel<<<blocks, TB_THREADS, TB_THREADS * 8 * sizeof(float), st>>>(reinterpret_cast<const uint4*>(x), npix, C, chunk, ws);
  PMU_LAUNCH_CHECK();
  f64_to_f32_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws, out, C);
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
  gauss_head_bwd_nhwc_kernel<<<B, TB_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(enc), w, dmu, dls,
                                                                        reinterpret_cast<__nv_bfloat16*>(denc), dw, db, C, h * w_, L);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
