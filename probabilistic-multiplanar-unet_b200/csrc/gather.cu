// K1 — plane slicing gather (data plane in).
//
// Replaces MRI_Dataset.sample_slice + preprocess (utils/mri_dataset.py:70-82,101-112):
//   plane 0: I[r,c] = vol[s,r,c]   plane 1: I[r,c] = vol[r,s,c]   plane 2: I[r,c] = vol[r,c,s]
// followed by I / max(I) when max(I) != 0 (fp64 divide, then .float()).
//
// All three kernels are HBM-bound integer-indexing/byte-moving work: the design rules
// are coalescing, 128-bit accesses, and enough independent loads in flight; plane 2
// (slice index on the fastest axis) goes through a shared-memory transpose tile so both
// the volume reads (along z) and the slice writes (along c) are full 128 B lines.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "pmu_common.cuh"
#include "sm100_ptx.cuh"

namespace pmu {

// ---------------------------------------------------------------------------------
// plane_max: per-slice maxima of all three planes in one pass over the volume.
// One warp per (x,y) row; lanes stride along z with float4.
// ---------------------------------------------------------------------------------
constexpr int PM_WARPS = 8;
constexpr int PM_ROWS_PER_WARP = 16;
constexpr int PM_ZSEG = 1024;  // z extent handled per pass: 32 lanes * 4 * 8 chunks

template <bool VEC>
__global__ void __launch_bounds__(PM_WARPS * 32)
plane_max_kernel(const float* __restrict__ vol, int d0, int d1, int d2, float* __restrict__ maxes) {
  __shared__ float zred[PM_WARPS][PM_ZSEG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nrows = (int64_t)d0 * d1;
  const int64_t row0 = ((int64_t)blockIdx.x * PM_WARPS + warp) * PM_ROWS_PER_WARP;
  float* max0 = maxes;
  float* max1 = maxes + d0;
  float* max2 = maxes + d0 + d1;
  for (int zseg = 0; zseg < d2; zseg += PM_ZSEG) {
    float zmax[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) zmax[j][i] = -INFINITY;
    for (int rr = 0; rr < PM_ROWS_PER_WARP; ++rr) {
      const int64_t row = row0 + rr;
      if (row >= nrows) break;
      const int x = (int)(row / d1), y = (int)(row % d1);
      const float* p = vol + row * d2 + zseg;
      float rmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int z = 4 * (lane + 32 * j);
        if (zseg + z < d2) {
          float v[4];
          if (VEC) {
            float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(p + z));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (zseg + z + i < d2) ? __ldg(p + z + i) : -INFINITY;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            zmax[j][i] = fmaxf(zmax[j][i], v[i]);
            rmax = fmaxf(rmax, v[i]);
          }
        }
      }
      rmax = warp_max(rmax);
      if (lane == 0) {
        if (rmax == 0.f) rmax = 0.f;  // canonicalise -0
        atomic_max_float(max0 + x, rmax);
        atomic_max_float(max1 + y, rmax);
      }
    }
    // combine the 8 warps' per-z maxima in shared memory: one atomic per z per block
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) zred[warp][4 * (lane + 32 * j) + i] = zmax[j][i];
    __syncthreads();
    for (int z = threadIdx.x; z < PM_ZSEG; z += PM_WARPS * 32) {
      float m = zred[0][z];
#pragma unroll
      for (int w = 1; w < PM_WARPS; ++w) m = fmaxf(m, zred[w][z]);
      if (zseg + z < d2 && m != -INFINITY) {
        if (m == 0.f) m = 0.f;
        atomic_max_float(max2 + zseg + z, m);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// exact gather, planes 0 and 1: row copies (contiguous along z == c).
//   grid.y = slice b, grid.x strides over the H*W/4 float4 of the slice.
// ---------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ vol, int d1, int d2, int plane, int s0, int H, int W,
                   const float* __restrict__ max_in, float* __restrict__ max_out,
                   float* __restrict__ out) {
  const int b = blockIdx.y, s = s0 + b;
  const float m = max_in ? __ldg(max_in + s) : 0.f;
  const int64_t hw = (int64_t)H * W;
  float* dst = out + (int64_t)b * hw;
  float tmax = -INFINITY;
  if (VEC) {
    const int w4 = W >> 2;
    const int n4 = (int)(hw >> 2);
    const int stride = gridDim.x * blockDim.x;
    // 4 independent 128-bit loads in flight per thread, then 4 stores
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * stride;
        if (i < n4) {
          const int r = i / w4, c4 = i - r * w4;
          // plane 0: vol[s][r][c]; plane 1: vol[r][s][c]
          const int64_t src = (plane == 0) ? ((int64_t)s * d1 + r) * d2 : ((int64_t)r * d1 + s) * d2;
          v[u] = ldg_stream_f4(reinterpret_cast<const float4*>(vol + src) + c4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * stride;
        if (i < n4) {
          float4 t = v[u];
          tmax = fmaxf(tmax, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
          if (max_in) {
            t.x = ref_normalise(t.x, m); t.y = ref_normalise(t.y, m);
            t.z = ref_normalise(t.z, m); t.w = ref_normalise(t.w, m);
          }
          stg_stream_f4(reinterpret_cast<float4*>(dst) + i, t);
        }
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
         i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / W), c = (int)(i % W);
      const int64_t src = (plane == 0) ? ((int64_t)s * d1 + r) * d2 + c : ((int64_t)r * d1 + s) * d2 + c;
      float v = __ldg(vol + src);
      tmax = fmaxf(tmax, v);
      dst[i] = max_in ? ref_normalise(v, m) : v;
    }
  }
  if (max_out) {
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0 && tmax != -INFINITY) {
      if (tmax == 0.f) tmax = 0.f;
      atomic_max_float(max_out + b, tmax);
    }
  }
}

// ---------------------------------------------------------------------------------
// exact gather, plane 2: out[b][r][c] = vol[r][c][s0+b].  For a fixed r this is a
// [c][s] -> [b][c] transpose; tile 64 (c) x 64 (s) through shared memory.
//   grid = (c tiles, s tiles, r)
// ---------------------------------------------------------------------------------
constexpr int T2_C = 64, T2_S = 64;

template <bool VEC>
__global__ void __launch_bounds__(256)
gather_plane2_kernel(const float* __restrict__ vol, int d1, int d2, int s0, int ns, int H, int W,
                     const float* __restrict__ max_in, float* __restrict__ max_out,
                     float* __restrict__ out) {
  __shared__ float tile[T2_C][T2_S + 1];
  const int r = blockIdx.z;
  const int c0 = blockIdx.x * T2_C, b0 = blockIdx.y * T2_S;
  const int t = threadIdx.x;
  // ---- read: 64 c-rows x 64 s; 16 float4 (or 64 scalars) per c-row; 4 loads in flight per thread ----
  if (VEC) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = t + 256 * k;  // 0..1023
      const int cc = idx >> 4, s4 = (idx & 15) << 2;
      const int c = c0 + cc, b = b0 + s4;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < W && b < ns)   // ns % 4 == 0 guaranteed on the VEC path
        v[k] = ldg_stream_f4(reinterpret_cast<const float4*>(vol + ((int64_t)r * d1 + c) * d2 + s0 + b));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = t + 256 * k;
      const int cc = idx >> 4, s4 = (idx & 15) << 2;
      tile[cc][s4 + 0] = v[k].x; tile[cc][s4 + 1] = v[k].y; tile[cc][s4 + 2] = v[k].z; tile[cc][s4 + 3] = v[k].w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int idx = t + 256 * k;  // 0..4095
      const int cc = idx >> 6, ss = idx & 63;
      const int c = c0 + cc, b = b0 + ss;
      tile[cc][ss] = (c < W && b < ns) ? __ldg(vol + ((int64_t)r * d1 + c) * d2 + s0 + b) : 0.f;
    }
  }
  __syncthreads();
  // ---- write: 64 s-rows x 64 c; 16 float4 per s-row ----
  const int64_t hw = (int64_t)H * W;
  if (VEC) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = t + 256 * k;
      const int ss = idx >> 4, c4 = (idx & 15) << 2;
      const int b = b0 + ss, c = c0 + c4;
      float4 v = make_float4(tile[c4 + 0][ss], tile[c4 + 1][ss], tile[c4 + 2][ss], tile[c4 + 3][ss]);
      const bool ok = (b < ns && c < W);  // W % 4 == 0 on the VEC path
      float tmax = ok ? fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) : -INFINITY;
      if (ok) {
        if (max_in) {
          const float m = __ldg(max_in + s0 + b);
          v.x = ref_normalise(v.x, m); v.y = ref_normalise(v.y, m);
          v.z = ref_normalise(v.z, m); v.w = ref_normalise(v.w, m);
        }
        stg_stream_f4(reinterpret_cast<float4*>(out + (int64_t)b * hw + (int64_t)r * W + c), v);
      }
      if (max_out) {  // 16 consecutive lanes share one slice b
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if ((t & 15) == 0 && tmax != -INFINITY) {
          if (tmax == 0.f) tmax = 0.f;
          atomic_max_float(max_out + b, tmax);
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int idx = t + 256 * k;
      const int ss = idx >> 6, cc = idx & 63;
      const int b = b0 + ss, c = c0 + cc;
      if (b < ns && c < W) {
        float v = tile[cc][ss];
        if (max_out) {
          float mv = (v == 0.f) ? 0.f : v;
          atomic_max_float(max_out + b, mv);
        }
        if (max_in) v = ref_normalise(v, __ldg(max_in + s0 + b));
        out[(int64_t)b * hw + (int64_t)r * W + c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// affine-grid gather: nearest / trilinear, zeros outside.  fp32 arithmetic in EXACTLY the op order of the oracle
// (oracle/resample_fma.c): q = fma(c, v, fma(r, u, fma(s, n, o))), lerp(a, b, t) = fma(t, b - a, a) — explicit
// __fmaf_rn / __fsub_rn / __fadd_rn, so nothing is left to the compiler's contraction and the result is bit-identical
// to oracle.resample_slices.  (Round 1 specified separately rounded multiplies and adds: 75 instructions per output
// pixel, which bounded the brick kernel at 0.23-0.26 of HBM; the fused spec needs half as many.)
// ---------------------------------------------------------------------------------
struct Affine12 { float a[12]; };

__device__ __forceinline__ float fetch_vox(const float* __restrict__ vol, int d0, int d1, int d2,
                                           int ix, int iy, int iz) {
  if (ix < 0 || ix >= d0 || iy < 0 || iy >= d1 || iz < 0 || iz >= d2) return 0.f;
  return __ldg(vol + ((int64_t)ix * d1 + iy) * d2 + iz);
}
__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
  return __fmaf_rn(t, __fsub_rn(b, a), a);
}

template <bool TRILINEAR>
__global__ void __launch_bounds__(256)
gather_affine_kernel(const float* __restrict__ vol, int d0, int d1, int d2, Affine12 A, int s0,
                     int H, int W, const float* __restrict__ max_in, float* __restrict__ max_out,
                     float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t hw = (int64_t)H * W;
  const float sf = (float)(s0 + b);
  const float m = max_in ? __ldg(max_in + s0 + b) : 0.f;
  float tmax = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float rf = (float)(int)(i / W), cf = (float)(int)(i % W);
    float q[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
      q[ax] = __fmaf_rn(cf, A.a[9 + ax], __fmaf_rn(rf, A.a[6 + ax], __fmaf_rn(sf, A.a[3 + ax], A.a[ax])));
    float v;
    if (!TRILINEAR) {
      const int ix = (int)floorf(__fadd_rn(q[0], 0.5f));
      const int iy = (int)floorf(__fadd_rn(q[1], 0.5f));
      const int iz = (int)floorf(__fadd_rn(q[2], 0.5f));
      v = fetch_vox(vol, d0, d1, d2, ix, iy, iz);
    } else {
      const float fx = floorf(q[0]), fy = floorf(q[1]), fz = floorf(q[2]);
      const float tx = __fsub_rn(q[0], fx), ty = __fsub_rn(q[1], fy), tz = __fsub_rn(q[2], fz);
      const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
      const float c00 = lerp_rn(fetch_vox(vol, d0, d1, d2, x0, y0, z0), fetch_vox(vol, d0, d1, d2, x0, y0, z0 + 1), tz);
      const float c01 = lerp_rn(fetch_vox(vol, d0, d1, d2, x0, y0 + 1, z0), fetch_vox(vol, d0, d1, d2, x0, y0 + 1, z0 + 1), tz);
      const float c10 = lerp_rn(fetch_vox(vol, d0, d1, d2, x0 + 1, y0, z0), fetch_vox(vol, d0, d1, d2, x0 + 1, y0, z0 + 1), tz);
      const float c11 = lerp_rn(fetch_vox(vol, d0, d1, d2, x0 + 1, y0 + 1, z0), fetch_vox(vol, d0, d1, d2, x0 + 1, y0 + 1, z0 + 1), tz);
      v = lerp_rn(lerp_rn(c00, c01, ty), lerp_rn(c10, c11, ty), tx);
    }
    tmax = fmaxf(tmax, v);
    out[(int64_t)b * hw + i] = max_in ? ref_normalise(v, m) : v;
  }
  if (max_out) {
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0 && tmax != -INFINITY) {
      if (tmax == 0.f) tmax = 0.f;
      atomic_max_float(max_out + b, tmax);
    }
  }
}

// ---------------------------------------------------------------------------------
// affine-grid gather with a TMA-staged volume brick: the block's 1024 output pixels (a ts x tr x tc
// tile of slices x rows x cols, chosen on the host so that the tile is long along the output axis
// that walks z) read their taps from a [BX][BY][BZ] brick of the volume that ONE 3-D TMA box load
// brings into shared memory — full-line reads along z, out-of-volume voxels zero-filled by TMA
// (== the zeros padding of the resampling spec).  Coordinates / interpolation are exactly those of
// gather_affine_kernel, so the result is bit-identical.
// ---------------------------------------------------------------------------------
struct BrickCfg { int ts, tr, tc; int lg_tc, lg_tr; int bx, by, bz; int tiles_r, tiles_c; };

__device__ __forceinline__ void affine_q(const Affine12& A, float sf, float rf, float cf, float (&q)[3]) {
#pragma unroll
  for (int ax = 0; ax < 3; ++ax)
    q[ax] = __fmaf_rn(cf, A.a[9 + ax], __fmaf_rn(rf, A.a[6 + ax], __fmaf_rn(sf, A.a[3 + ax], A.a[ax])));
}

template <bool TRILINEAR>
__global__ void __launch_bounds__(256)
gather_affine_brick_kernel(const __grid_constant__ CUtensorMap tmV, Affine12 A, BrickCfg g, int s0, int ns, int H, int W,
                           const float* __restrict__ max_in, float* __restrict__ max_out, float* __restrict__ out) {
  extern __shared__ __align__(128) float brick[];     // [bx][by][bz]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int org[3];
  __shared__ float smax[32];
  const int t = threadIdx.x;
  const int tile_c = blockIdx.x % g.tiles_c, tile_r = blockIdx.x / g.tiles_c;
  const int sl0 = blockIdx.y * g.ts, r0 = tile_r * g.tr, c0 = tile_c * g.tc;
  const uint32_t bar = ptx::smem_u32(&mbar);
  if (t < 32) smax[t] = -INFINITY;
  if (t == 0) {
    // brick origin = floor of the minimum tap coordinate over the tile's 8 corners (affine: extremes at corners)
    float mn[3] = {INFINITY, INFINITY, INFINITY};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int sl = min(sl0 + ((k & 4) ? g.ts - 1 : 0), ns - 1), r = min(r0 + ((k & 2) ? g.tr - 1 : 0), H - 1),
                c = min(c0 + ((k & 1) ? g.tc - 1 : 0), W - 1);
      float q[3];
      affine_q(A, (float)(s0 + sl), (float)r, (float)c, q);
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) mn[ax] = fminf(mn[ax], TRILINEAR ? floorf(q[ax]) : floorf(__fadd_rn(q[ax], 0.5f)));
    }
    org[0] = (int)mn[0]; org[1] = (int)mn[1];
    org[2] = (int)mn[2] & ~3;    // TMA wants the innermost start 16-byte aligned (the box is 4 floats longer for it)
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)(g.bx * g.by * g.bz * 4));
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(brick)), "l"(reinterpret_cast<uint64_t>(&tmV)), "r"(bar), "r"(org[2]), "r"(org[1]), "r"(org[0])
        : "memory");
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);
  const int ox = org[0], oy = org[1], oz = org[2];
  const int sy = g.bz, sx = g.by * g.bz;          // brick strides
  auto tap = [&](int ix, int iy, int iz) -> float {
    return brick[(ix - ox) * sx + (iy - oy) * sy + (iz - oz)];
  };
  // 4 output pixels per thread, pixel = t + 256 k: the 32 lanes of a warp are 32 CONSECUTIVE output columns (tc >= 32:
  // one row of one slice), so on a grid whose columns walk z their taps fall into 32 consecutive shared-memory banks and
  // their stores into one 128-byte line.  (A thread owning 4 consecutive columns strides the lanes by 16 B: 4-way bank
  // conflicts on all 8 taps put the shared-memory pipe at 83 % and the kernel at 0.26 of HBM, profiles/r02_gather_oblique.txt.)
  const int K0 = -(ox * sx + oy * sy + oz);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = t + 256 * k;                  // 0 .. ts*tr*tc - 1 (== 1023)
    const int cl = idx & (g.tc - 1), rl = (idx >> g.lg_tc) & (g.tr - 1), sl = idx >> (g.lg_tc + g.lg_tr);   // powers of two
    const int b = sl0 + sl, r = r0 + rl, c = c0 + cl;
    if (!((b < ns) && (r < H) && (c < W))) continue;
    float q[3];
    affine_q(A, (float)(s0 + b), (float)r, (float)c, q);
    float v;
    if (!TRILINEAR) {
      const int ix = __float2int_rd(__fadd_rn(q[0], 0.5f)), iy = __float2int_rd(__fadd_rn(q[1], 0.5f)), iz = __float2int_rd(__fadd_rn(q[2], 0.5f));
      v = brick[ix * sx + iy * sy + iz + K0];
    } else {
      // floor as one conversion each way (cvt.rmi.s32.f32, cvt.rn.f32.s32): the integer floor is exact for |q| < 2^24
      const int x0 = __float2int_rd(q[0]), y0 = __float2int_rd(q[1]), z0 = __float2int_rd(q[2]);
      const float tx = __fsub_rn(q[0], (float)x0), ty = __fsub_rn(q[1], (float)y0), tz = __fsub_rn(q[2], (float)z0);
      const float* bp = brick + (x0 * sx + y0 * sy + z0 + K0);   // one base, 8 offsets
      const float c00 = lerp_rn(bp[0], bp[1], tz);
      const float c01 = lerp_rn(bp[sy], bp[sy + 1], tz);
      const float c10 = lerp_rn(bp[sx], bp[sx + 1], tz);
      const float c11 = lerp_rn(bp[sx + sy], bp[sx + sy + 1], tz);
      v = lerp_rn(lerp_rn(c00, c01, ty), lerp_rn(c10, c11, ty), tx);
    }
    if (max_out) { float mv = (v == 0.f) ? 0.f : v; atomic_max_float(&smax[sl], mv); }
    const float m = max_in ? __ldg(max_in + s0 + b) : 0.f;
    out[((int64_t)b * H + r) * W + c] = max_in ? ref_normalise(v, m) : v;
  }
  if (max_out) {
    __syncthreads();
    if (t < g.ts && sl0 + t < ns && smax[t] != -INFINITY) atomic_max_float(max_out + sl0 + t, smax[t]);
  }
}

__global__ void __launch_bounds__(256)
slice_normalize_kernel(float* __restrict__ slices, const float* __restrict__ slice_max, int64_t hw) {
  const int b = blockIdx.y;
  const float m = __ldg(slice_max + b);
  if (m == 0.f) return;
  float* p = slices + (int64_t)b * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = ref_normalise(p[i], m);
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

}  // namespace pmu

using namespace pmu;

extern "C" int pmu_fill_f32(float* p, float value, int64_t n, void* stream) {
  PMU_CHECK_ARG(p != nullptr && n >= 0, "pmu_fill_f32: bad arguments");
  if (n == 0) return PMU_OK;
  const int blocks = (int)std::min<int64_t>(cdiv64(n, 256), (int64_t)sm_count() * 8);
  fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, value, n);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_plane_max(const float* vol, const int32_t dims[3], float* maxes, void* stream) {
  PMU_CHECK_ARG(vol && dims && maxes, "pmu_plane_max: null pointer");
  const int d0 = dims[0], d1 = dims[1], d2 = dims[2];
  PMU_CHECK_ARG(d0 > 0 && d1 > 0 && d2 > 0, "pmu_plane_max: dims must be positive");
  const int64_t nrows = (int64_t)d0 * d1;
  const int64_t blocks = cdiv64(nrows, PM_WARPS * PM_ROWS_PER_WARP);
  const bool vec = (d2 % 4 == 0) && aligned16(vol);
  if (vec)
    plane_max_kernel<true><<<(unsigned)blocks, PM_WARPS * 32, 0, (cudaStream_t)stream>>>(vol, d0, d1, d2, maxes);
  else
    plane_max_kernel<false><<<(unsigned)blocks, PM_WARPS * 32, 0, (cudaStream_t)stream>>>(vol, d0, d1, d2, maxes);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_slice_gather(const float* vol, const int32_t dims[3], int plane, int s0, int ns,
                                int interp, const float* affine_host, int H, int W,
                                const float* slice_max_in, float* slice_max_out, float* out,
                                void* stream) {
  PMU_CHECK_ARG(ns >= 0, "pmu_slice_gather: negative slice count");
  if (ns == 0) return PMU_OK;  // empty range: nothing to do (out may be a zero-size buffer)
  PMU_CHECK_ARG(vol && dims && out, "pmu_slice_gather: null pointer");
  const int d0 = dims[0], d1 = dims[1], d2 = dims[2];
  PMU_CHECK_ARG(d0 > 0 && d1 > 0 && d2 > 0, "pmu_slice_gather: dims must be positive");
  PMU_CHECK_ARG(plane >= 0 && plane <= 2, "pmu_slice_gather: plane must be 0, 1 or 2 (got %d)", plane);
  PMU_CHECK_ARG(ns >= 0 && s0 >= 0 && H > 0 && W > 0, "pmu_slice_gather: bad slice range / size");
  if (ns == 0) return PMU_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (interp == PMU_INTERP_EXACT) {
    const int ext[3] = {d0, d1, d2};
    PMU_CHECK_ARG(s0 + ns <= ext[plane], "pmu_slice_gather: slices [%d,%d) exceed extent %d of plane %d",
                  s0, s0 + ns, ext[plane], plane);
    const int eh = (plane == 0) ? d1 : d0, ew = (plane == 2) ? d1 : d2;
    PMU_CHECK_ARG(H == eh && W == ew, "pmu_slice_gather: EXACT needs H,W = %d,%d (got %d,%d)", eh, ew, H, W);
    if (plane < 2) {
      const bool vec = (d2 % 4 == 0) && aligned16(vol) && aligned16(out);
      const int64_t work = vec ? ((int64_t)H * W / 4) : (int64_t)H * W;
      dim3 grid((unsigned)std::min<int64_t>(cdiv64(work, 256 * 4), 4096), ns);   // 4 vectors per thread
      if (vec)
        gather_rows_kernel<true><<<grid, 256, 0, st>>>(vol, d1, d2, plane, s0, H, W, slice_max_in, slice_max_out, out);
      else
        gather_rows_kernel<false><<<grid, 256, 0, st>>>(vol, d1, d2, plane, s0, H, W, slice_max_in, slice_max_out, out);
    } else {
      PMU_CHECK_ARG(H <= 65535, "pmu_slice_gather: plane 2 supports H <= 65535");
      const bool vec = (d2 % 4 == 0) && (s0 % 4 == 0) && (ns % 4 == 0) && (W % 4 == 0) &&
                       aligned16(vol) && aligned16(out);
      dim3 grid(cdiv(W, T2_C), cdiv(ns, T2_S), H);
      if (vec)
        gather_plane2_kernel<true><<<grid, 256, 0, st>>>(vol, d1, d2, s0, ns, H, W, slice_max_in, slice_max_out, out);
      else
        gather_plane2_kernel<false><<<grid, 256, 0, st>>>(vol, d1, d2, s0, ns, H, W, slice_max_in, slice_max_out, out);
    }
  } else if (interp == PMU_INTERP_NEAREST || interp == PMU_INTERP_TRILINEAR) {
    PMU_CHECK_ARG(affine_host != nullptr, "pmu_slice_gather: affine grid needs 12 host floats");
    Affine12 A;
    for (int i = 0; i < 12; ++i) A.a[i] = affine_host[i];
    // Fast path: on a standard plane grid (o = 0, n/u/v = the view's unit vectors) every tap weight is exactly
    // 0 or 1, so nearest and trilinear reproduce plain slicing bit for bit (oracle test_oracle_slicing_identities,
    // GPU test_gather_affine_bit_exact) — hand the launch to the HBM-rate copy kernels.
    {
      static const float std_aff[3][12] = {{0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1},
                                           {0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1},
                                           {0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0}};
      for (int pl = 0; pl < 3; ++pl) {
        bool same = true;
        for (int i = 0; i < 12; ++i) same = same && (A.a[i] == std_aff[pl][i]);
        const int ext3[3] = {d0, d1, d2};
        const int eh = (pl == 0) ? d1 : d0, ew = (pl == 2) ? d1 : d2;
        if (same && H == eh && W == ew && s0 + ns <= ext3[pl])
          return pmu_slice_gather(vol, dims, pl, s0, ns, PMU_INTERP_EXACT, nullptr, H, W, slice_max_in, slice_max_out, out, stream);
      }
    }
    // ---- TMA-staged brick path ----
    if (d2 % 4 == 0 && aligned16(vol)) {
      // Tile = ts x tr x tc output pixels (slices x rows x columns, powers of two, 1024 per block, tc >= 32: a warp is 32
      // consecutive output columns).  The brick the tile's taps touch is (tile extent along each volume axis + 3)
      // voxels: a flat 1 x 32 x 32 tile of an oblique grid reads 3 x-planes to produce one — 4x read amplification,
      // which is what held this kernel at 0.23 of HBM.  Pick the tile shape with the smallest brick.
      BrickCfg g = {};
      double best = 1e30;
      for (int lts = 0; lts <= 5; ++lts)
        for (int ltr = 0; lts + ltr <= 8; ++ltr) {
          const int ltc = 10 - lts - ltr;
          if (ltc < 5 || ltc > 8) continue;      // a warp = 32 consecutive output columns
          const int ts = 1 << lts, tr = 1 << ltr, tc = 1 << ltc;
          int e3[3];
          for (int ax = 0; ax < 3; ++ax)
            e3[ax] = (int)floorf((ts - 1) * fabsf(A.a[3 + ax]) + (tr - 1) * fabsf(A.a[6 + ax]) + (tc - 1) * fabsf(A.a[9 + ax])) + 3;
          const double vol_b = (double)e3[0] * e3[1] * (((e3[2] + 3) & ~3) + 4);
          // prefer tiles that are long along the output columns (16-byte stores) when the bricks are equal
          const double cost = vol_b - 1e-3 * tc;
          if (cost < best) { best = cost; g.ts = ts; g.tr = tr; g.tc = tc; g.lg_tr = ltr; g.lg_tc = ltc; }
        }
      int ext[3];
      for (int ax = 0; ax < 3; ++ax) {
        const float e = (g.ts - 1) * fabsf(A.a[3 + ax]) + (g.tr - 1) * fabsf(A.a[6 + ax]) + (g.tc - 1) * fabsf(A.a[9 + ax]);
        ext[ax] = (int)floorf(e) + 3;     // floor offset + second tap + rounding slack
      }
      g.bx = ext[0]; g.by = ext[1]; g.bz = ((ext[2] + 3) & ~3) + 4;   // inner extent: multiple of 16 B, + slack for the aligned start
      const size_t bytes = (size_t)g.bx * g.by * g.bz * 4;
      static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
      if (!enc) {
        void* fp = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
      }
      int cc_major = 0, dev = 0;
      PMU_CUDA(cudaGetDevice(&dev));
      PMU_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
      // (a box larger than the tensor extent is not a valid TMA box: tiny volumes use the plain kernel)
      if (enc && cc_major >= 9 && bytes <= 96 * 1024 && g.bx <= std::min(256, d0) && g.by <= std::min(256, d1) &&
          g.bz <= std::min(256, d2) && cdiv(ns, g.ts) <= 65535) {
        CUtensorMap tm;
        cuuint64_t dims3[3] = {(cuuint64_t)d2, (cuuint64_t)d1, (cuuint64_t)d0};
        cuuint64_t str3[2] = {(cuuint64_t)d2 * 4, (cuuint64_t)d1 * d2 * 4};
        cuuint32_t box3[3] = {(cuuint32_t)g.bz, (cuuint32_t)g.by, (cuuint32_t)g.bx};
        cuuint32_t es3[3] = {1, 1, 1};
        CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(vol), dims3, str3, box3, es3,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr == CUDA_SUCCESS) {
          g.tiles_r = cdiv(H, g.tr); g.tiles_c = cdiv(W, g.tc);
          dim3 grid(g.tiles_r * g.tiles_c, cdiv(ns, g.ts));
          if (interp == PMU_INTERP_TRILINEAR) {
            PMU_CUDA(cudaFuncSetAttribute(gather_affine_brick_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            gather_affine_brick_kernel<true><<<grid, 256, bytes, st>>>(tm, A, g, s0, ns, H, W, slice_max_in, slice_max_out, out);
          } else {
            PMU_CUDA(cudaFuncSetAttribute(gather_affine_brick_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            gather_affine_brick_kernel<false><<<grid, 256, bytes, st>>>(tm, A, g, s0, ns, H, W, slice_max_in, slice_max_out, out);
          }
          PMU_LAUNCH_CHECK();
          return PMU_OK;
        }
      }
    }
    dim3 grid((unsigned)std::min<int64_t>(cdiv64((int64_t)H * W, 256), 2048), ns);
    if (interp == PMU_INTERP_TRILINEAR)
      gather_affine_kernel<true><<<grid, 256, 0, st>>>(vol, d0, d1, d2, A, s0, H, W, slice_max_in, slice_max_out, out);
    else
      gather_affine_kernel<false><<<grid, 256, 0, st>>>(vol, d0, d1, d2, A, s0, H, W, slice_max_in, slice_max_out, out);
  } else {
    PMU_CHECK_ARG(false, "pmu_slice_gather: unknown interp %d", interp);
  }
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_slice_normalize(float* slices, const float* slice_max, int ns, int64_t hw, void* stream) {
  PMU_CHECK_ARG(slices && slice_max && ns >= 0 && hw > 0, "pmu_slice_normalize: bad arguments");
  if (ns == 0) return PMU_OK;
  dim3 grid((unsigned)std::min<int64_t>(cdiv64(hw, 256), 1024), ns);
  slice_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slices, slice_max, hw);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
