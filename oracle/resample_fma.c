/* CPU oracle of the build-defined slice resampling — TEST INFRASTRUCTURE ONLY, like the rest of
 * oracle/: compiled with gcc by oracle/pmu_oracle.py (and by __graft_entry__.build()) into oracle/_build/, loaded with
 * ctypes, called by tests/ only.  The reference has no resampling (its three views are plain numpy slicing,
 * utils/mri_dataset.py:70-82); this spec generalises them to an affine slice grid and reproduces them bit for bit on the
 * identity grids.
 *
 * Spec (every operation a correctly rounded fp32 operation; C here because numpy has no fused multiply-add):
 *   q[ax]   = fma(c, v[ax], fma(r, u[ax], fma(s, n[ax], o[ax])))          output pixel (s, r, c) -> voxel coordinate
 *   nearest : index = floor(q + 0.5) per axis
 *   trilinear: 8 taps around floor(q), t = q - floor(q); lerp(a, b, t) = fma(t, b - a, a) along z, then y, then x
 *   taps outside the volume read 0 (zeros padding).
 * The CUDA kernels (csrc/gather.cu) use __fmaf_rn / __fsub_rn / __fadd_rn in exactly this order. */
#include <math.h>
#include <stdint.h>

static float fetch(const float* vol, int d0, int d1, int d2, long ix, long iy, long iz) {
  if (ix < 0 || ix >= d0 || iy < 0 || iy >= d1 || iz < 0 || iz >= d2) return 0.0f;
  return vol[((int64_t)ix * d1 + iy) * d2 + iz];
}
static float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }

void pmu_oracle_resample(const float* vol, int d0, int d1, int d2, const float* aff, int s0, int ns, int H, int W,
                         int trilinear, float* out) {
  for (int b = 0; b < ns; ++b)
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        float q[3];
        for (int ax = 0; ax < 3; ++ax)
          q[ax] = fmaf((float)c, aff[9 + ax], fmaf((float)r, aff[6 + ax], fmaf((float)(s0 + b), aff[3 + ax], aff[ax])));
        float v;
        if (!trilinear) {
          v = fetch(vol, d0, d1, d2, (long)floorf(q[0] + 0.5f), (long)floorf(q[1] + 0.5f), (long)floorf(q[2] + 0.5f));
        } else {
          const float fx = floorf(q[0]), fy = floorf(q[1]), fz = floorf(q[2]);
          const float tx = q[0] - fx, ty = q[1] - fy, tz = q[2] - fz;
          const long x0 = (long)fx, y0 = (long)fy, z0 = (long)fz;
          const float c00 = lerp(fetch(vol, d0, d1, d2, x0, y0, z0), fetch(vol, d0, d1, d2, x0, y0, z0 + 1), tz);
          const float c01 = lerp(fetch(vol, d0, d1, d2, x0, y0 + 1, z0), fetch(vol, d0, d1, d2, x0, y0 + 1, z0 + 1), tz);
          const float c10 = lerp(fetch(vol, d0, d1, d2, x0 + 1, y0, z0), fetch(vol, d0, d1, d2, x0 + 1, y0, z0 + 1), tz);
          const float c11 = lerp(fetch(vol, d0, d1, d2, x0 + 1, y0 + 1, z0), fetch(vol, d0, d1, d2, x0 + 1, y0 + 1, z0 + 1), tz);
          v = lerp(lerp(c00, c01, ty), lerp(c10, c11, ty), tx);
        }
        out[((int64_t)b * H + r) * W + c] = v;
      }
}

/* Nearest-voxel scatter of per-slice values back onto the lattice, with a per-voxel count (SURVEY.md App. A step 6 for
 * non-identity grids): voxel = floor(q + 0.5) of the output pixel; pixels that land outside the volume are dropped.
 * vals [ns][K][H][W] (K channels per pixel), acc [d0][K][d1][d2] += vals, cnt [d0][d1][d2] += 1. */
void pmu_oracle_scatter_nearest(const float* vals, int K, const float* aff, int s0, int ns, int H, int W, int d0, int d1,
                                int d2, float* acc, float* cnt) {
  for (int b = 0; b < ns; ++b)
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        long i[3];
        for (int ax = 0; ax < 3; ++ax)
          i[ax] = (long)floorf(fmaf((float)c, aff[9 + ax], fmaf((float)r, aff[6 + ax], fmaf((float)(s0 + b), aff[3 + ax], aff[ax]))) + 0.5f);
        if (i[0] < 0 || i[0] >= d0 || i[1] < 0 || i[1] >= d1 || i[2] < 0 || i[2] >= d2) continue;
        for (int k = 0; k < K; ++k)
          acc[(((int64_t)i[0] * K + k) * d1 + i[1]) * d2 + i[2]] += vals[(((int64_t)b * K + k) * H + r) * W + c];
        cnt[((int64_t)i[0] * d1 + i[1]) * d2 + i[2]] += 1.0f;
      }
}
