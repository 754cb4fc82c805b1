// bf16 NHWC helper layers around the tcgen05 convolution (all HBM-bound):
//   first conv (Cin = 1 or 2: a CUDA-core stencil, arithmetic intensity ~9 flop/B)
//   2x2 pooling (MaxPool2d(2), unet_parts.py:33; AvgPool2d(2,2,ceil_mode), probabilistic_unet.py:36)
//   NHWC bf16 -> NCHW fp32 (hands unet_features back in the reference's layout)
#include "pmu_common.cuh"
#include "h16.cuh"

namespace pmu {

// ---------------------------------------------------------------------------------
// first layer: x fp32 [B][Cin][H][W] (Cin = 1, or 2 given as two 1-channel tensors)
//   -> y bf16 [B][H][W][Cout].  thread = (pixel, group of 8 couts); a warp covers 4
//   consecutive pixels x 64 couts = 512 contiguous output bytes.
// ---------------------------------------------------------------------------------
constexpr int FL_MAX_W = 128 * 2 * 9;  // Cout <= 128, Cin <= 2

template <bool F16>
__global__ void __launch_bounds__(256)
conv3x3_first_bf16_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                          const float* __restrict__ w, const float* __restrict__ bias,
                          __nv_bfloat16* __restrict__ y, int B, int H, int W, int Cin, int Cout, int relu) {
  __shared__ __align__(16) float w_s[FL_MAX_W];  // [ci*9+tap][Cout]
  __shared__ float b_s[128];
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += 256) {
    const int co = i / (Cin * 9), rem = i % (Cin * 9);
    w_s[rem * Cout + co] = __ldg(w + i);
  }
  for (int i = threadIdx.x; i < Cout; i += 256) b_s[i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int groups = Cout >> 3;
  const int64_t total = (int64_t)B * H * W * groups;
  const int64_t HW = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int b = (int)(pix / HW);
    const int64_t rem = pix % HW;
    const int h = (int)(rem / W), ww = (int)(rem % W);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = b_s[g * 8 + c];
    for (int ci = 0; ci < Cin; ++ci) {
      const float* src = (ci == 0 ? x0 : x1) + (int64_t)b * HW;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = h + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ww + kx - 1;
          const float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(src + (int64_t)iy * W + ix) : 0.f;
          const float* wp = w_s + (ci * 9 + ky * 3 + kx) * Cout + g * 8;
          const float4 wa = *reinterpret_cast<const float4*>(wp);
          const float4 wb = *reinterpret_cast<const float4*>(wp + 4);
          acc[0] = fmaf(v, wa.x, acc[0]); acc[1] = fmaf(v, wa.y, acc[1]);
          acc[2] = fmaf(v, wa.z, acc[2]); acc[3] = fmaf(v, wa.w, acc[3]);
          acc[4] = fmaf(v, wb.x, acc[4]); acc[5] = fmaf(v, wb.y, acc[5]);
          acc[6] = fmaf(v, wb.z, acc[6]); acc[7] = fmaf(v, wb.w, acc[7]);
        }
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = acc[2 * j], c = acc[2 * j + 1];
      if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
      pk[j] = pack16_rn<F16>(a, c);
    }
    *reinterpret_cast<uint4*>(y + pix * Cout + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Fast path (Cin == 1, W % 4 == 0 — every slice of the hot path): thread = 4 consecutive pixels
// x 8 couts.  Weights sit in shared memory as [tap][Cout] so a thread's 8 weights of a tap are two
// broadcast LDS.128 (18 per 288 FMA); ~80 registers -> 3 blocks of 256 threads per SM hide the
// load latency.  Inputs come in as one float4 + two halo scalars per row; 32-bit index math only.
// Per 4 pixels: 9 global loads, 288 FMA, 4 x 16-byte stores (a warp writes four full 128-byte
// lines per store instruction).
template <bool F16>
__global__ void __launch_bounds__(256, 3)
conv3x3_first_c1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                        int Cout, int relu) {
  __shared__ __align__(16) float w_s[9 * 128];   // [tap][Cout]
  __shared__ __align__(16) float b_s[128];
  for (int i = threadIdx.x; i < Cout * 9; i += 256) w_s[(i % 9) * Cout + i / 9] = __ldg(w + i);
  for (int i = threadIdx.x; i < Cout; i += 256) b_s[i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int groups = Cout >> 3;
  const int g = threadIdx.x % groups;                     // cout group (fastest: 8 lanes share a pixel quad)
  const int qpb = 256 / groups;                           // pixel quads per block
  const int W4 = W >> 2;
  const unsigned total = (unsigned)B * H * W4;            // pixel quads
  const float4 ba = *reinterpret_cast<const float4*>(b_s + g * 8), bb = *reinterpret_cast<const float4*>(b_s + g * 8 + 4);
  for (unsigned q = blockIdx.x * qpb + threadIdx.x / groups; q < total; q += gridDim.x * qpb) {
    const int w0 = (int)(q % W4) * 4;
    const unsigned row = q / W4;                          // b*H + h
    const int h = (int)(row % H);
    const float* src = x + (size_t)row * W + w0;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      acc[p][0] = ba.x; acc[p][1] = ba.y; acc[p][2] = ba.z; acc[p][3] = ba.w;
      acc[p][4] = bb.x; acc[p][5] = bb.y; acc[p][6] = bb.z; acc[p][7] = bb.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = h + ky - 1;
      if (iy < 0 || iy >= H) continue;
      const float* rp = src + (ky - 1) * W;
      const float4 mid = __ldg(reinterpret_cast<const float4*>(rp));
      const float lft = (w0 > 0) ? __ldg(rp - 1) : 0.f;
      const float rgt = (w0 + 4 < W) ? __ldg(rp + 4) : 0.f;
      const float iv[6] = {lft, mid.x, mid.y, mid.z, mid.w, rgt};
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wp = w_s + (ky * 3 + kx) * Cout + g * 8;
        const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(iv[p + kx], wv[c], acc[p][c]);
      }
    }
    __nv_bfloat16* dst = y + ((size_t)row * W + w0) * Cout + g * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = acc[p][2 * j], c = acc[p][2 * j + 1];
        if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
        pk[j] = pack16_rn<F16>(a, c);
      }
      *reinterpret_cast<uint4*>(dst + (size_t)p * Cout) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// ---------------------------------------------------------------------------------
// 2x2 stride-2 pooling, NHWC bf16; thread = (output pixel, 8 channels = 16 B)
// ---------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256)
pool2_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                  int C, int Ho, int Wo, int mode) {
  const int groups = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * groups;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int g = (int)(i % groups);
    int64_t r = i / groups;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float m[8], s[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { m[c] = -INFINITY; s[c] = 0.f; }
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < H && ix < W) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + iy) * W + ix) * C + g * 8));
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f2 = unpack16<F16>(u[j]);
            const float lo = f2.x, hi = f2.y;
            m[2 * j] = fmaxf(m[2 * j], lo); m[2 * j + 1] = fmaxf(m[2 * j + 1], hi);
            s[2 * j] += lo; s[2 * j + 1] += hi;
          }
          ++cnt;
        }
      }
    const float inv = 1.f / (float)cnt;
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = (mode == PMU_POOL_MAX) ? m[2 * j] : s[2 * j] * inv;
      const float c = (mode == PMU_POOL_MAX) ? m[2 * j + 1] : s[2 * j + 1] * inv;
      pk[j] = pack16_rn<F16>(a, c);
    }
    *reinterpret_cast<uint4*>(y + (((int64_t)b * Ho + oy) * Wo + ox) * C + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------------------------
// [B][HW][C] bf16 -> [B][C][HW] fp32 through a 64 px x 64 ch shared tile
// ---------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t HW, int C) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  // read: 64 px rows, 64 channels each (32 x bf16x2 per row)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = t + 256 * k;  // 0..2047
    const int pp = idx >> 5, c2 = (idx & 31) * 2;
    float lo = 0.f, hi = 0.f;
    if (p0 + pp < HW && c0 + c2 < C) {
      const float2 f2 = unpack16<F16>(*reinterpret_cast<const uint32_t*>(x + ((int64_t)b * HW + p0 + pp) * C + c0 + c2));
      lo = f2.x; hi = f2.y;
    }
    tile[pp][c2] = lo; tile[pp][c2 + 1] = hi;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = t + 256 * k;  // 0..4095
    const int cc = idx >> 6, pp = idx & 63;
    if (p0 + pp < HW && c0 + cc < C) y[((int64_t)b * C + c0 + cc) * HW + p0 + pp] = tile[pp][cc];
  }
}

// ---------------------------------------------------------------------------------
// [B][C][HW] fp32 -> [B][HW][C] bf16 through a 64 px x 64 ch shared tile (training: the tensor-core
// convolutions of the bf16 training mode take NHWC bf16 operands; BatchNorm etc. stay fp32 NCHW)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t HW, int C) {
  __shared__ float tile[64][65];     // [channel][pixel]
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = t + 256 * k;  // 0..4095
    const int cc = idx >> 6, pp = idx & 63;
    tile[cc][pp] = (p0 + pp < HW && c0 + cc < C) ? __ldg(x + ((int64_t)b * C + c0 + cc) * HW + p0 + pp) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = t + 256 * k;  // 0..2047
    const int pp = idx >> 5, c2 = (idx & 31) * 2;
    if (p0 + pp < HW && c0 + c2 < C)
      *reinterpret_cast<__nv_bfloat162*>(y + ((int64_t)b * HW + p0 + pp) * C + c0 + c2) =
          __floats2bfloat162_rn(tile[c2][pp], tile[c2 + 1][pp]);
  }
}

}  // namespace pmu

namespace pmu {
int conv_first_tc_launch(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int relu,
                         int f16, cudaStream_t st);   // conv_tc.cu
}
using namespace pmu;

extern "C" int pmu_conv3x3_first_bf16(const float* x0, const float* x1, const float* w, const float* bias,
                                      void* y, int B, int H, int W, int Cin, int Cout, int relu, int f16, void* stream) {
  PMU_CHECK_ARG(x0 && w && y && B > 0 && H > 0 && W > 0, "pmu_conv3x3_first_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(Cin >= 1 && Cin <= 2 && (Cin == 1 || x1), "pmu_conv3x3_first_bf16: Cin must be 1, or 2 with x1 (got %d)", Cin);
  PMU_CHECK_SUPPORTED(Cout % 8 == 0 && Cout <= 128, "pmu_conv3x3_first_bf16: Cout must be a multiple of 8, <= 128 (got %d)", Cout);
  PMU_CHECK_ARG(aligned16(y), "pmu_conv3x3_first_bf16: y must be 16-byte aligned");
  // tensor-core path (conv_tc.cu): im2col rows in shared memory, two K = 16 UMMAs per 128 pixels, TMA-store epilogue
  if (Cin == 1 && Cout == 64) {
    const int rc = conv_first_tc_launch(x0, w, bias, y, B, H, W, relu, f16, (cudaStream_t)stream);
    if (rc != PMU_ERR_UNSUPPORTED) return rc;
  }
  const int groups = Cout / 8;
  if (Cin == 1 && W % 4 == 0 && 256 % groups == 0 && aligned16(x0) && (int64_t)B * H * (W / 4) < (1ll << 31)) {
    const int64_t quads = (int64_t)B * H * (W / 4);
    const int qpb = 256 / groups;
    const int blocks = (int)std::min<int64_t>(cdiv64(quads, qpb), (int64_t)sm_count() * 16);
    if (f16) conv3x3_first_c1_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x0, w, bias, reinterpret_cast<__nv_bfloat16*>(y), B, H, W, Cout, relu);
    else conv3x3_first_c1_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x0, w, bias, reinterpret_cast<__nv_bfloat16*>(y), B, H, W, Cout, relu);
    PMU_LAUNCH_CHECK();
    return PMU_OK;
  }
  const int64_t total = (int64_t)B * H * W * (Cout / 8);
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 32);
  if (f16) conv3x3_first_bf16_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x0, x1, w, bias, reinterpret_cast<__nv_bfloat16*>(y), B, H, W, Cin, Cout, relu);
  else conv3x3_first_bf16_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x0, x1, w, bias, reinterpret_cast<__nv_bfloat16*>(y), B, H, W, Cin, Cout, relu);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_pool2_bf16(const void* x, void* y, int B, int H, int W, int C, int mode, int f16, void* stream) {
  PMU_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0, "pmu_pool2_bf16: bad arguments");
  PMU_CHECK_ARG(mode == PMU_POOL_MAX || mode == PMU_POOL_AVG_CEIL, "pmu_pool2_bf16: unknown mode %d", mode);
  PMU_CHECK_SUPPORTED(C % 8 == 0, "pmu_pool2_bf16: C must be a multiple of 8 (got %d)", C);
  PMU_CHECK_ARG(aligned16(x) && aligned16(y), "pmu_pool2_bf16: pointers must be 16-byte aligned");
  const int Ho = (mode == PMU_POOL_MAX) ? H / 2 : (H + 1) / 2;
  const int Wo = (mode == PMU_POOL_MAX) ? W / 2 : (W + 1) / 2;
  PMU_CHECK_ARG(Ho > 0 && Wo > 0, "pmu_pool2_bf16: input too small");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 32);
  if (f16) pool2_bf16_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C, Ho, Wo, mode);
  else pool2_bf16_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C, Ho, Wo, mode);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int C, int f16, void* stream) {
  PMU_CHECK_ARG(x && y && B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0, "pmu_nhwc_bf16_to_nchw_f32: bad arguments");
  PMU_CHECK_SUPPORTED(C % 2 == 0, "pmu_nhwc_bf16_to_nchw_f32: C must be even");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)cdiv64(HW, 64), cdiv(C, 64), B);
  if (f16) nhwc_to_nchw_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, HW, C);
  else nhwc_to_nchw_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, HW, C);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

// space-to-depth of a bf16 NHWC map: dst[b,h,w,(i*2+j)*C + c] = src[b,2h+i,2w+j,c]  (16-byte vectors; C % 8 == 0).
// The gradient of a k2 s2 transposed convolution reads its output-side tensor exactly in this order, which turns its
// data and weight gradients into plain 1x1 GEMMs with K (resp. N) = 4*C (nn.ConvTranspose2d backward, unet_parts.py:52).
namespace pmu {
__global__ void __launch_bounds__(256)
s2d_nhwc_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int C8, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C8);
    int64_t r = i / C8;
    const int ij = (int)(r & 3); r >>= 2;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int64_t b = r / H;
    const int64_t sp = (b * 2 * H + 2 * h + (ij >> 1)) * (2 * (int64_t)W) + 2 * w + (ij & 1);
    dst[i] = __ldg(src + sp * C8 + c);
  }
}
}  // namespace pmu

extern "C" int pmu_s2d_nhwc_bf16(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  PMU_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0, "pmu_s2d_nhwc_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(C % 8 == 0 && aligned16(x) && aligned16(y), "pmu_s2d_nhwc_bf16: C must be a multiple of 8, pointers 16-byte aligned");
  const int64_t total = (int64_t)B * H * W * 4 * (C / 8);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 32);
  pmu::s2d_nhwc_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y),
                                                                   H, W, C / 8, total);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int H, int W, int C, void* stream) {
  PMU_CHECK_ARG(x && y && B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0, "pmu_nchw_f32_to_nhwc_bf16: bad arguments");
  PMU_CHECK_SUPPORTED(C % 2 == 0, "pmu_nchw_f32_to_nhwc_bf16: C must be even");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)cdiv64(HW, 64), cdiv(C, 64), B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), HW, C);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
