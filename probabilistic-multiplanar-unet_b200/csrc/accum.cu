// K4 — softmax + scatter-accumulate + finalise (data plane out), and K5 — the small
// reductions of the training step / evaluation.
//
// Replaces eval.py:157 (softmax over classes), eval.py:176-190 (torch.cat + permute of the
// per-view probability volumes onto [x,C,y,z]), eval.py:193 (fusion) and extends it with the
// build-defined N-sample sum / sum-of-squares, variance and entropy.
// All of it is HBM-bound: float4 accesses, coalesced along z (the contiguous voxel axis);
// plane 2 (slice index == z) goes through a shared-memory transpose tile.
#include "pmu_common.cuh"

namespace pmu {

constexpr int MAXC = 8;

// ---------------------------------------------------------------------------------
// unfused softmax + accumulate over N samples: logits [B][N][C][HW] -> sums [B][2][C][HW]
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_accum_kernel(const float* __restrict__ logits, float* __restrict__ sums, int N, int C, int64_t HW) {
  const int b = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  float s1[MAXC], s2[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) s1[c] = s2[c] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* lp = logits + ((int64_t)b * N + n) * C * HW + p;
    float lg[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      lg[c] = (c < C) ? __ldg(lp + (int64_t)c * HW) : -INFINITY;
      mx = fmaxf(mx, lg[c]);
    }
    float e[MAXC], den = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      e[c] = (c < C) ? expf(lg[c] - mx) : 0.f;
      den += e[c];
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const float pr = e[c] * inv;
      s1[c] += pr;
      s2[c] = fmaf(pr, pr, s2[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c < C) {
      sums[(((int64_t)b * 2 + 0) * C + c) * HW + p] = s1[c];
      sums[(((int64_t)b * 2 + 1) * C + c) * HW + p] = s2[c];
    }
}

// ---------------------------------------------------------------------------------
// scatter-accumulate, planes 0 and 1 (rows stay contiguous along z == c).
//   sums [ns][2][C][H][W];  S1,S2 [X][C][Y][Z].   grid = (chunks, ns)
// ---------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ sums, int plane, int s0, int C, int H, int W, int Y, int Z,
                    float* __restrict__ S1, float* __restrict__ S2) {
  const int b = blockIdx.y, s = s0 + b;
  const int64_t hw = (int64_t)H * W, chw = (int64_t)C * hw;
  const float* src1 = sums + (int64_t)b * 2 * chw;
  const float* src2 = src1 + chw;
  constexpr int V = VEC ? 4 : 1;
  const int64_t n = chw / V;
  const int wv = W / V;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % wv) * V;
    const int r = (int)((i / wv) % H);
    const int k = (int)(i / ((int64_t)wv * H));
    // plane 0: voxel (s, r, c);  plane 1: voxel (r, s, c)
    const int64_t dst = (plane == 0) ? (((int64_t)s * C + k) * Y + r) * Z + c
                                     : (((int64_t)r * C + k) * Y + s) * Z + c;
    if (VEC) {
      const float4 a1 = ldg_stream_f4(reinterpret_cast<const float4*>(src1) + i);
      const float4 a2 = ldg_stream_f4(reinterpret_cast<const float4*>(src2) + i);
      float4 d1 = *reinterpret_cast<float4*>(S1 + dst);
      float4 d2 = *reinterpret_cast<float4*>(S2 + dst);
      d1.x += a1.x; d1.y += a1.y; d1.z += a1.z; d1.w += a1.w;
      d2.x += a2.x; d2.y += a2.y; d2.z += a2.z; d2.w += a2.w;
      *reinterpret_cast<float4*>(S1 + dst) = d1;
      *reinterpret_cast<float4*>(S2 + dst) = d2;
    } else {
      S1[dst] += __ldg(src1 + i);
      S2[dst] += __ldg(src2 + i);
    }
  }
}

// ---------------------------------------------------------------------------------
// scatter-accumulate, plane 2: voxel (r, c, s).  For fixed (j,k,r): [b][c] -> [c][s].
//   grid = (c tiles, s tiles, 2*C*H)
// ---------------------------------------------------------------------------------
constexpr int S2_C = 64, S2_S = 32;

template <bool VEC>
__global__ void __launch_bounds__(256)
scatter_plane2_kernel(const float* __restrict__ sums, int s0, int ns, int C, int H, int W, int Y, int Z,
                      float* __restrict__ S1, float* __restrict__ S2) {
  __shared__ float tile[S2_S][S2_C + 1];
  const int t = threadIdx.x;
  const int r = blockIdx.z % H, k = (blockIdx.z / H) % C, j = blockIdx.z / (H * C);
  const int c0 = blockIdx.x * S2_C, b0 = blockIdx.y * S2_S;
  const int64_t hw = (int64_t)H * W;
  // read 32 b-rows x 64 c
  if (VEC) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = t + 256 * q;
      const int bb = idx >> 4, c4 = (idx & 15) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + bb < ns && c0 + c4 < W)
        v = ldg_stream_f4(reinterpret_cast<const float4*>(
            sums + (((int64_t)(b0 + bb) * 2 + j) * C + k) * hw + (int64_t)r * W + c0 + c4));
      tile[bb][c4 + 0] = v.x; tile[bb][c4 + 1] = v.y; tile[bb][c4 + 2] = v.z; tile[bb][c4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = t + 256 * q;
      const int bb = idx >> 6, cc = idx & 63;
      tile[bb][cc] = (b0 + bb < ns && c0 + cc < W)
                         ? __ldg(sums + (((int64_t)(b0 + bb) * 2 + j) * C + k) * hw + (int64_t)r * W + c0 + cc)
                         : 0.f;
    }
  }
  __syncthreads();
  float* S = j ? S2 : S1;
  // write 64 c-rows x 32 s (contiguous along z)
  if (VEC) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = t + 256 * q;
      const int cc = idx >> 3, s4 = (idx & 7) << 2;
      if (c0 + cc < W && b0 + s4 < ns) {
        float* dp = S + (((int64_t)r * C + k) * Y + (c0 + cc)) * Z + s0 + b0 + s4;
        float4 d = *reinterpret_cast<float4*>(dp);
        d.x += tile[s4 + 0][cc]; d.y += tile[s4 + 1][cc]; d.z += tile[s4 + 2][cc]; d.w += tile[s4 + 3][cc];
        *reinterpret_cast<float4*>(dp) = d;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = t + 256 * q;
      const int cc = idx >> 5, ss = idx & 31;
      if (c0 + cc < W && b0 + ss < ns)
        S[(((int64_t)r * C + k) * Y + (c0 + cc)) * Z + s0 + b0 + ss] += tile[ss][cc];
    }
  }
}

// ---------------------------------------------------------------------------------
// [build-defined] scatter for a NON-identity slice grid: every output pixel adds its sums to
// the voxel NEAREST to its coordinate q = fma(c, v, fma(r, u, fma(s, n, o))) (the resampling gather's coordinate, same
// op order as oracle/resample_fma.c), pixels outside the volume are dropped, and a per-voxel weight counts the
// contributions.  Pixels of different slices can land on one voxel: fp32 atomics (red.global.add).
// ---------------------------------------------------------------------------------
struct ScAffine { float a[12]; };
__global__ void __launch_bounds__(256)
scatter_affine_kernel(const float* __restrict__ sums, ScAffine A, int s0, int ns, int H, int W, int d0, int d1, int d2,
                      int C, float weight, float* __restrict__ S1, float* __restrict__ S2, float* __restrict__ cnt) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)ns * hw;
  const int64_t yz = (int64_t)d1 * d2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int b = (int)(i / hw);
    const int64_t p = i % hw;
    const float sf = (float)(s0 + b), rf = (float)(int)(p / W), cf = (float)(int)(p % W);
    int v3[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const float q = __fmaf_rn(cf, A.a[9 + ax], __fmaf_rn(rf, A.a[6 + ax], __fmaf_rn(sf, A.a[3 + ax], A.a[ax])));
      v3[ax] = __float2int_rd(__fadd_rn(q, 0.5f));
    }
    if (v3[0] < 0 || v3[0] >= d0 || v3[1] < 0 || v3[1] >= d1 || v3[2] < 0 || v3[2] >= d2) continue;
    const int64_t vox = (int64_t)v3[1] * d2 + v3[2];
    const float* src = sums + (int64_t)b * 2 * C * hw + p;
    for (int k = 0; k < C; ++k) {
      atomicAdd(S1 + ((int64_t)v3[0] * C + k) * yz + vox, __ldg(src + (int64_t)k * hw));
      atomicAdd(S2 + ((int64_t)v3[0] * C + k) * yz + vox, __ldg(src + (int64_t)(C + k) * hw));
    }
    atomicAdd(cnt + (int64_t)v3[0] * yz + vox, weight);
  }
}

// finalise with a per-voxel count: mean = S1 / cnt, var = max(S2 / cnt - mean^2, 0); voxels nothing landed on read 0
__global__ void __launch_bounds__(256)
finalize_counted_kernel(const float* __restrict__ S1, const float* __restrict__ S2, const float* __restrict__ cnt, int64_t X,
                        int C, int64_t YZ, float* __restrict__ mean, float* __restrict__ var, float* __restrict__ entropy,
                        uint8_t* __restrict__ labels) {
  const int64_t total = X * YZ;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t x = i / YZ, q = i % YZ;
    const float n = __ldg(cnt + i);
    const float inv = (n > 0.f) ? 1.f / n : 0.f;
    float ent = 0.f, best = -INFINITY;
    int lab = 0;
    for (int k = 0; k < C; ++k) {
      const int64_t off = (x * C + k) * YZ + q;
      const float m = __ldg(S1 + off) * inv;
      if (mean) mean[off] = m;
      if (var) var[off] = fmaxf(__ldg(S2 + off) * inv - m * m, 0.f);
      ent -= (m > 0.f) ? m * logf(m) : 0.f;
      if (m > best) { best = m; lab = k; }
    }
    if (entropy) entropy[i] = ent;
    if (labels) labels[i] = (uint8_t)lab;
  }
}

// ---------------------------------------------------------------------------------
// finalise: mean / variance / entropy / argmax labels.  Thread = V consecutive (y,z).
// ---------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ S1, const float* __restrict__ S2, float inv_count, int64_t X,
                int C, int64_t YZ, float* __restrict__ mean, float* __restrict__ var,
                float* __restrict__ entropy, uint8_t* __restrict__ labels) {
  constexpr int V = VEC ? 4 : 1;
  const int64_t nyz = YZ / V, total = X * nyz;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t x = i / nyz, q = (i % nyz) * V;
    float ent[V], best[V];
    int lab[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { ent[v] = 0.f; best[v] = -INFINITY; lab[v] = 0; }
    for (int k = 0; k < C; ++k) {
      const int64_t off = (x * C + k) * YZ + q;
      float a[V], b2[V];
      if (VEC) {
        const float4 t1 = ldg_stream_f4(reinterpret_cast<const float4*>(S1 + off));
        a[0] = t1.x; a[1 % V] = t1.y; a[2 % V] = t1.z; a[3 % V] = t1.w;
        if (var) {
          const float4 t2 = ldg_stream_f4(reinterpret_cast<const float4*>(S2 + off));
          b2[0] = t2.x; b2[1 % V] = t2.y; b2[2 % V] = t2.z; b2[3 % V] = t2.w;
        }
      } else {
        a[0] = __ldg(S1 + off);
        if (var) b2[0] = __ldg(S2 + off);
      }
      float m[V], vr[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        m[v] = a[v] * inv_count;
        vr[v] = var ? fmaxf(b2[v] * inv_count - m[v] * m[v], 0.f) : 0.f;
        ent[v] -= (m[v] > 0.f) ? m[v] * logf(m[v]) : 0.f;
        if (m[v] > best[v]) { best[v] = m[v]; lab[v] = k; }
      }
      if (VEC) {
        if (mean) stg_stream_f4(reinterpret_cast<float4*>(mean + off), make_float4(m[0], m[1 % V], m[2 % V], m[3 % V]));
        if (var) stg_stream_f4(reinterpret_cast<float4*>(var + off), make_float4(vr[0], vr[1 % V], vr[2 % V], vr[3 % V]));
      } else {
        if (mean) mean[off] = m[0];
        if (var) var[off] = vr[0];
      }
    }
    const int64_t o = x * YZ + q;
    if (VEC) {
      if (entropy) stg_stream_f4(reinterpret_cast<float4*>(entropy + o), make_float4(ent[0], ent[1 % V], ent[2 % V], ent[3 % V]));
      if (labels) {
        uchar4 l = make_uchar4((unsigned char)lab[0], (unsigned char)lab[1 % V], (unsigned char)lab[2 % V], (unsigned char)lab[3 % V]);
        *reinterpret_cast<uchar4*>(labels + o) = l;
      }
    } else {
      if (entropy) entropy[o] = ent[0];
      if (labels) labels[o] = (uint8_t)lab[0];
    }
  }
}

// ---------------------------------------------------------------------------------
// K5 reductions.  Block partials are warp-shuffle reduced, one atomicAdd per block.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  return t;  // valid in thread 0
}

// CE(reduction none) summed over batch and pixels (probabilistic_unet.py:288,303-304)
__global__ void __launch_bounds__(256)
ce_sum_kernel(const float* __restrict__ logits, const float* __restrict__ target, int C, int64_t HW,
              int64_t total, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / HW, p = i % HW;
    const float* lp = logits + b * C * HW + p;
    const float tf = __ldg(target + i);
    const int t = (int)tf;                 // .long() truncation of the float label
    // nn.CrossEntropyLoss raises on a target outside [0, C) ("Target k is out of bounds"): counted here, raised by the caller
    if (!(tf > -1.f && tf < (float)C)) ++bad;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(lp + (int64_t)c * HW));
    float den = 0.f, lt = 0.f;
    for (int c = 0; c < C; ++c) {
      const float l = __ldg(lp + (int64_t)c * HW);
      den += expf(l - mx);
      if (c == t) lt = l;
    }
    acc += (logf(den) + mx) - lt;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(out + 1, 1.f);     // out[1] > 0: some label was out of range
}

// analytic KL(q||p), diagonal Gaussians (probabilistic_unet.py:272)
__global__ void kl_kernel(const float* __restrict__ mu_q, const float* __restrict__ ls_q,
                          const float* __restrict__ mu_p, const float* __restrict__ ls_p, int B, int L,
                          float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int l = 0; l < L; ++l) {
    const int i = b * L + l;
    const float sq = expf(ls_q[i]), sp = expf(ls_p[i]);
    const float ratio = sq / sp, vr = ratio * ratio;
    const float d = (mu_q[i] - mu_p[i]) / sp;
    s += 0.5f * (vr + d * d - 1.f - logf(vr));
  }
  out[b] = s;
}

// dice_loss.py:10-12 — three global sums
__global__ void __launch_bounds__(256)
dice_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                 float* __restrict__ sums) {
  __shared__ float red[8];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float p = __ldg(pred + i), t = __ldg(target + i);
    a = fmaf(p, t, a); b += p; c += t;
  }
  a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
  if (threadIdx.x == 0) { atomicAdd(sums + 0, a); atomicAdd(sums + 1, b); atomicAdd(sums + 2, c); }
}

// eval.py:42-49 — one-hot(argmax) vs (truth == k), k = 1..C-1, on the [X][C][YZ] layout
__global__ void __launch_bounds__(256)
argmax_dice_kernel(const float* __restrict__ prob, const float* __restrict__ truth, int64_t X, int C,
                   int64_t YZ, float* __restrict__ sums) {
  __shared__ float red[8];
  float acc[(MAXC - 1) * 3];
#pragma unroll
  for (int i = 0; i < (MAXC - 1) * 3; ++i) acc[i] = 0.f;
  const int64_t total = X * YZ;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t x = i / YZ, q = i % YZ;
    float best = -INFINITY;
    int lab = 0;
    for (int k = 0; k < C; ++k) {
      const float v = __ldg(prob + (x * C + k) * YZ + q);
      if (v > best) { best = v; lab = k; }
    }
    const float tr = __ldg(truth + i);
#pragma unroll
    for (int k = 1; k < MAXC; ++k) {
      if (k < C) {
        const float pk = (lab == k) ? 1.f : 0.f, tk = (tr == (float)k) ? 1.f : 0.f;
        acc[(k - 1) * 3 + 0] += pk * tk;
        acc[(k - 1) * 3 + 1] += pk;
        acc[(k - 1) * 3 + 2] += tk;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < (MAXC - 1) * 3; ++i) {
    if (i < (C - 1) * 3) {
      const float s = block_sum(acc[i], red);
      if (threadIdx.x == 0) atomicAdd(sums + i, s);
    }
  }
}

}  // namespace pmu

using namespace pmu;

extern "C" int pmu_softmax_accum(const float* logits, float* slice_sums, int B, int N, int C, int64_t HW,
                                 void* stream) {
  PMU_CHECK_ARG(logits && slice_sums && B > 0 && B <= 65535 && N > 0 && HW > 0, "pmu_softmax_accum: bad arguments");
  PMU_CHECK_SUPPORTED(C > 0 && C <= MAXC, "pmu_softmax_accum: C must be in 1..8 (got %d)", C);
  dim3 grid((unsigned)cdiv64(HW, 256), B);
  softmax_accum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, slice_sums, N, C, HW);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_scatter_accum(const float* slice_sums, int plane, int s0, int ns, const int32_t dims[3],
                                 int C, float* S1, float* S2, void* stream) {
  PMU_CHECK_ARG(slice_sums && dims && S1 && S2, "pmu_scatter_accum: null pointer");
  PMU_CHECK_ARG(plane >= 0 && plane <= 2, "pmu_scatter_accum: plane must be 0..2 (got %d)", plane);
  const int X = dims[0], Y = dims[1], Z = dims[2];
  const int ext[3] = {X, Y, Z};
  PMU_CHECK_ARG(X > 0 && Y > 0 && Z > 0 && C > 0, "pmu_scatter_accum: bad dims");
  PMU_CHECK_ARG(s0 >= 0 && ns >= 0 && s0 + ns <= ext[plane], "pmu_scatter_accum: slices [%d,%d) exceed extent %d",
                s0, s0 + ns, ext[plane]);
  if (ns == 0) return PMU_OK;
  const int H = (plane == 0) ? Y : X, W = (plane == 2) ? Y : Z;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = aligned16(slice_sums) && aligned16(S1) && aligned16(S2);
  if (plane < 2) {
    const bool vec = al && (Z % 4 == 0);
    const int64_t work = (int64_t)C * H * W / (vec ? 4 : 1);
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(work, 256 * 2), 4096), ns);
    if (vec) scatter_rows_kernel<true><<<grid, 256, 0, st>>>(slice_sums, plane, s0, C, H, W, Y, Z, S1, S2);
    else scatter_rows_kernel<false><<<grid, 256, 0, st>>>(slice_sums, plane, s0, C, H, W, Y, Z, S1, S2);
  } else {
    PMU_CHECK_ARG((int64_t)2 * C * H <= 65535, "pmu_scatter_accum: 2*C*H must be <= 65535 for plane 2");
    const bool vec = al && (Z % 4 == 0) && (W % 4 == 0) && (s0 % 4 == 0) && (ns % 4 == 0);
    dim3 grid(cdiv(W, S2_C), cdiv(ns, S2_S), 2 * C * H);
    if (vec) scatter_plane2_kernel<true><<<grid, 256, 0, st>>>(slice_sums, s0, ns, C, H, W, Y, Z, S1, S2);
    else scatter_plane2_kernel<false><<<grid, 256, 0, st>>>(slice_sums, s0, ns, C, H, W, Y, Z, S1, S2);
  }
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fuse_finalize(const float* S1, const float* S2, float count, const int32_t dims[3], int C,
                                 float* mean, float* var, float* entropy, uint8_t* labels, void* stream) {
  PMU_CHECK_ARG(S1 && dims && count > 0.f, "pmu_fuse_finalize: bad arguments");
  PMU_CHECK_ARG(!var || S2, "pmu_fuse_finalize: variance needs S2");
  PMU_CHECK_SUPPORTED(C > 0 && C <= 255, "pmu_fuse_finalize: C must be in 1..255");
  const int64_t X = dims[0], YZ = (int64_t)dims[1] * dims[2];
  PMU_CHECK_ARG(X > 0 && YZ > 0, "pmu_fuse_finalize: bad dims");
  const bool vec = (YZ % 4 == 0) && aligned16(S1) && (!S2 || aligned16(S2)) && (!mean || aligned16(mean)) &&
                   (!var || aligned16(var)) && (!entropy || aligned16(entropy)) &&
                   (!labels || (reinterpret_cast<uintptr_t>(labels) & 3u) == 0);
  const int64_t total = X * YZ / (vec ? 4 : 1);
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 16);
  const float inv = 1.0f / count;
  if (vec) finalize_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(S1, S2, inv, X, C, YZ, mean, var, entropy, labels);
  else finalize_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(S1, S2, inv, X, C, YZ, mean, var, entropy, labels);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_scatter_accum_affine(const float* slice_sums, const float* affine_host, int s0, int ns, int H, int W,
                                        const int32_t dims[3], int C, float weight, float* S1, float* S2, float* cnt,
                                        void* stream) {
  PMU_CHECK_ARG(ns >= 0, "pmu_scatter_accum_affine: negative slice count");
  if (ns == 0) return PMU_OK;
  PMU_CHECK_ARG(slice_sums && affine_host && dims && S1 && S2 && cnt, "pmu_scatter_accum_affine: null pointer");
  PMU_CHECK_ARG(H > 0 && W > 0 && C > 0 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && weight > 0.f, "pmu_scatter_accum_affine: bad shape");
  ScAffine A;
  for (int i = 0; i < 12; ++i) A.a[i] = affine_host[i];
  const int64_t total = (int64_t)ns * H * W;
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 16);
  scatter_affine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(slice_sums, A, s0, ns, H, W, dims[0], dims[1], dims[2], C, weight,
                                                                 S1, S2, cnt);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_fuse_finalize_counted(const float* S1, const float* S2, const float* cnt, const int32_t dims[3], int C,
                                         float* mean, float* var, float* entropy, uint8_t* labels, void* stream) {
  PMU_CHECK_ARG(S1 && cnt && dims, "pmu_fuse_finalize_counted: bad arguments");
  PMU_CHECK_ARG(!var || S2, "pmu_fuse_finalize_counted: variance needs S2");
  PMU_CHECK_SUPPORTED(C > 0 && C <= 255, "pmu_fuse_finalize_counted: C must be in 1..255");
  const int64_t X = dims[0], YZ = (int64_t)dims[1] * dims[2];
  PMU_CHECK_ARG(X > 0 && YZ > 0, "pmu_fuse_finalize_counted: bad dims");
  const int blocks = (int)std::min<int64_t>(cdiv64(X * YZ, 256), (int64_t)sm_count() * 16);
  finalize_counted_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(S1, S2, cnt, X, C, YZ, mean, var, entropy, labels);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_ce_sum(const float* logits, const float* target, int B, int C, int64_t HW, float* out,
                          void* stream) {
  PMU_CHECK_ARG(logits && target && out && B > 0 && C > 0 && HW > 0, "pmu_ce_sum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(float), st));
  const int64_t total = (int64_t)B * HW;
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)sm_count() * 8);
  ce_sum_kernel<<<blocks, 256, 0, st>>>(logits, target, C, HW, total, out);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_kl_diag_gauss(const float* mu_q, const float* log_sigma_q, const float* mu_p,
                                 const float* log_sigma_p, int B, int L, float* out, void* stream) {
  PMU_CHECK_ARG(mu_q && log_sigma_q && mu_p && log_sigma_p && out && B > 0 && L > 0, "pmu_kl_diag_gauss: bad arguments");
  kl_kernel<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(mu_q, log_sigma_q, mu_p, log_sigma_p, B, L, out);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_dice_sums(const float* pred, const float* target, int64_t n, float* sums, void* stream) {
  PMU_CHECK_ARG(pred && target && sums && n >= 0, "pmu_dice_sums: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), st));
  if (n == 0) return PMU_OK;
  const int blocks = (int)std::min<int64_t>(cdiv64(n, 256), (int64_t)sm_count() * 8);
  dice_sums_kernel<<<blocks, 256, 0, st>>>(pred, target, n, sums);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}

extern "C" int pmu_argmax_dice_sums(const float* prob, const float* truth, int64_t X, int C, int64_t YZ,
                                    float* sums, void* stream) {
  PMU_CHECK_ARG(prob && truth && sums && X > 0 && YZ > 0, "pmu_argmax_dice_sums: bad arguments");
  PMU_CHECK_SUPPORTED(C >= 2 && C <= MAXC, "pmu_argmax_dice_sums: C must be in 2..8 (got %d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  PMU_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * (C - 1) * 3, st));
  const int blocks = (int)std::min<int64_t>(cdiv64(X * YZ, 256), (int64_t)sm_count() * 8);
  argmax_dice_kernel<<<blocks, 256, 0, st>>>(prob, truth, X, C, YZ, sums);
  PMU_LAUNCH_CHECK();
  return PMU_OK;
}
