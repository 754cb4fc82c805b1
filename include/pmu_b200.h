/*
 * pmu_b200.h — C-ABI of the B200-native multi-planar probabilistic inference path.
 *
 * The reference (qzs634/Probabilistic-Multiplanar-Unet) is pure Python and has NO
 * FFI / plugin interface; its only seam is the Python class API in
 * model/probabilistic_unet/probabilistic_unet.py, model/unet/*.py, dice_loss.py and
 * the data plane in utils/mri_dataset.py / eval.py.  This header is the boundary the
 * build defines underneath that API: each entry point states the reference call
 * site (file:line, relative to Probabilistic-Multiplanar-Unet/) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - every function returns int: 0 = OK, negative = error (pmu_last_error() gives a
 *     thread-local message).  Nothing throws or aborts across the boundary.
 *   - the CALLER owns every buffer (device pointers unless marked host); the library
 *     allocates nothing on the device.
 *   - every launch is asynchronous on the caller-supplied cudaStream_t (passed as
 *     void*; NULL = legacy default stream).  Safe to capture into a CUDA graph.
 *   - device pointers must be 16-byte aligned (128-bit vector paths, TMA).
 *   - fp32 tensors of the "f32" family are NCHW (the reference's layout); bf16
 *     tensors of the "bf16" family are NHWC (channels-last, TMA/tcgen05-friendly).
 *   - volumes are fp32 [d0][d1][d2] = [x][y][z], z fastest (numpy C order of the
 *     array nibabel returns, mri_dataset.py:124).
 *   - voxel accumulators / outputs use the reference's avg_volume layout [x][C][y][z]
 *     (eval.py:176,193).
 */
#ifndef PMU_B200_H_
#define PMU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMU_OK 0
#define PMU_ERR_INVALID (-1)   /* bad argument (shape, alignment, enum) */
#define PMU_ERR_CUDA (-2)      /* CUDA runtime / driver error */
#define PMU_ERR_UNSUPPORTED (-3) /* shape outside what the sm_100a kernel supports */

#define PMU_INTERP_EXACT 0     /* axis-aligned integer slicing (mri_dataset.py:70-82) */
#define PMU_INTERP_NEAREST 1   /* affine grid, floor(q+0.5), zeros outside */
#define PMU_INTERP_TRILINEAR 2 /* affine grid, 8-tap lerp z,y,x, zeros outside */

#define PMU_DTYPE_F32 0
#define PMU_DTYPE_BF16 1

#define PMU_POOL_MAX 0         /* nn.MaxPool2d(2)                 unet_parts.py:33 */
#define PMU_POOL_AVG_CEIL 1    /* nn.AvgPool2d(2,2,0,ceil_mode)   probabilistic_unet.py:36 */

/* rows of the fp32 workspace of the two-level per-channel reductions on bf16 NHWC tensors (training step): each block of
 * the first pass writes one row of per-channel partials, the finalize pass adds the rows in fp64 (deterministic, no atomics) */
#define PMU_RED_MAX_BLOCKS 592

/* ---- housekeeping ------------------------------------------------------- */
const char* pmu_last_error(void);
int pmu_version(void);
/* sm count and compute capability of the current device. */
int pmu_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Select the CUDA device for this thread's subsequent calls (the library carries its own
 * statically linked CUDA runtime, so the caller's cudaSetDevice does not reach it).  The
 * stream passed to every call must belong to this device. */
int pmu_set_device(int device);

/* ---- launch context ("the library allocates nothing except inside an opaque pmu_ctx* ... with explicit
 * create/destroy").  The reference has no counterpart: a torch module keeps this state inside cuDNN / ATen handles
 * (model/unet/unet_parts.py:15-20 -> at::cudnn_convolution).  A context caches, per device, what a tcgen05 launch needs
 * besides its arguments: SM count and compute capability, the kernels' dynamic-shared-memory attributes and the TMA
 * descriptors (cuTensorMapEncodeTiled results keyed by pointer / extents / strides / box), so that the launch path makes
 * no driver query and encodes no descriptor in steady state.  pmu_ctx_bind makes a context current for the calling thread
 * (and selects its device); every entry point below then uses it.  NULL unbinds: each launch queries and encodes what it
 * needs, as pmu_set_device callers get.  A context may be bound by one thread at a time. */
typedef struct pmu_ctx pmu_ctx;
int pmu_ctx_create(int device, pmu_ctx** out);
int pmu_ctx_destroy(pmu_ctx* ctx);
int pmu_ctx_bind(pmu_ctx* ctx);
/* cached descriptors, cache hits and misses so far (any pointer may be NULL) */
int pmu_ctx_stats(pmu_ctx* ctx, int64_t* tensor_maps, int64_t* hits, int64_t* misses);

/* ---- K1: plane slicing (data plane in) ---------------------------------- *
 * replaces MRI_Dataset.sample_slice + preprocess, mri_dataset.py:70-82,101-112 */

/* Per-slice maxima of all three planes in ONE pass over the volume:
 * maxes[0..d0) plane 0, [d0..d0+d1) plane 1, [d0+d1..d0+d1+d2) plane 2.
 * (np.max(img_trans), mri_dataset.py:109).  maxes must be pre-filled by the caller
 * with -inf (pmu_fill_f32) — the kernel combines with atomic max. */
int pmu_plane_max(const float* vol, const int32_t dims[3], float* maxes, void* stream);

/* Gather ns slices [s0, s0+ns) of `plane` into out[ns][H][W] (C=1, so NCHW==NHWC).
 *  interp EXACT: H,W must equal the two remaining extents; bit-exact copies.
 *  interp NEAREST/TRILINEAR: affine = 12 HOST floats [o(3), n(3), u(3), v(3)],
 *    q = o + s*n + r*u + c*v in fp32 (fixed op order, see oracle resample_slices).
 *  slice_max_in  (nullable, indexed by absolute slice s): fuse the normalisation
 *    x / max if max != 0, computed with ONE IEEE fp32 division (div.rn.f32), which equals
 *    the reference's fp64 divide + .float() (mri_dataset.py:110,142) bit for bit: the fp64
 *    quotient of two fp32 values rounds to fp32 exactly like the correctly rounded fp32
 *    quotient (double rounding is innocuous for division with >= 2p+2 = 50 wide bits).
 *  slice_max_out (nullable, indexed by s - s0, pre-filled with -inf): records the
 *    max of each gathered (raw) slice with atomic max, for pmu_slice_normalize.
 *  out_dtype: PMU_DTYPE_F32 only (bf16 conversion happens in the first conv). */
int pmu_slice_gather(const float* vol, const int32_t dims[3], int plane, int s0, int ns,
                     int interp, const float* affine_host, int H, int W,
                     const float* slice_max_in, float* slice_max_out,
                     float* out, void* stream);

/* In-place x/max per slice (the same div.rn.f32), slices [ns][hw], slice_max [ns]:
 * MRI_Dataset.preprocess, mri_dataset.py:109-110 (x / np.max(x) if the max is not 0). */
int pmu_slice_normalize(float* slices, const float* slice_max, int ns, int64_t hw, void* stream);

/* p[0..n) = value (the -inf pre-fill of the max buffers above; no reference counterpart). */
int pmu_fill_f32(float* p, float value, int64_t n, void* stream);

/* ---- fp32 NCHW layer ops (parity mode, CUDA cores) ----------------------- */

/* y = [relu](conv3x3_pad1(cat(x0[B,C0,H,W], x1[B,C1,H,W]), w[Cout,C0+C1,3,3]) + bias).
 * x1 may be NULL (C1 = 0).  BatchNorm is folded into w/bias by the caller.
 * replaces nn.Conv2d+BatchNorm2d+ReLU (unet_parts.py:15-20, probabilistic_unet.py:38-45)
 * and the F.pad/torch.cat of Up.forward (unet_parts.py:58-66: skip first, up second). */
int pmu_conv3x3_f32(const float* x0, int C0, const float* x1, int C1, const float* w,
                    const float* bias, float* y, int B, int H, int W, int Cout, int relu,
                    void* stream);
/* y[B,Cout,H,W] = [relu](conv1x1(x[B,Cin,H,W], w[Cout,Cin]) + bias)   (unet_parts.py:73) */
int pmu_conv1x1_f32(const float* x, const float* w, const float* bias, float* y, int B,
                    int Cin, int Cout, int64_t HW, int relu, void* stream);
/* nn.ConvTranspose2d(k=2,s=2) (unet_parts.py:52): x[B,Cin,H,W], w[Cin,Cout,2,2] ->
 * y[B,Cout,Ho,Wo] placed at offset (padT,padL) inside a zero-filled [Ho,Wo] canvas
 * (the F.pad of unet_parts.py:61-62; Ho>=2H, Wo>=2W). */
int pmu_convt2x2_f32(const float* x, const float* w, const float* bias, float* y, int B,
                     int Cin, int Cout, int H, int W, int Ho, int Wo, int padT, int padL,
                     void* stream);
/* 2x2 stride-2 pooling; MAX floors the output size (nn.MaxPool2d(2), unet_parts.py:33), AVG_CEIL ceils it and
 * divides by the number of in-bounds taps (nn.AvgPool2d(2, 2, 0, ceil_mode=True), probabilistic_unet.py:36). */
int pmu_pool2_f32(const float* x, float* y, int B, int C, int H, int W, int mode, void* stream);
/* AxisAlignedConvGaussian head (probabilistic_unet.py:97-108): mean over H then W of
 * enc[B,C,h,w], 1x1 conv w[2L,C]+b -> mu[B,L], log_sigma[B,L]. */
int pmu_gauss_head_f32(const float* enc, const float* w, const float* b, float* mu,
                       float* log_sigma, int B, int C, int h, int w_, int L, void* stream);

/* Fcomb (probabilistic_unet.py:155-181) for N latent samples per slice.
 *  feat[B,F,H,W] fp32; z[B,N,L]; w0[F,F+L], b0[F]; wmid[(nl-2)][F,F], bmid[(nl-2)][F]
 *  (nl = no_convs_fcomb >= 2); wlast[C,F], blast[C].
 *  logits (nullable) [B,N,C,H,W]; slice_sums (nullable) [B,2,C,H,W] receives
 *  sum_n softmax and sum_n softmax^2 (App. A steps 5-6; overwritten, not added). */
int pmu_fcomb_f32(const float* feat, const float* z, const float* w0, const float* b0,
                  const float* wmid, const float* bmid, const float* wlast, const float* blast,
                  float* logits, float* slice_sums, int B, int N, int F, int L, int C,
                  int nl, int64_t HW, void* stream);

/* ---- 16-bit NHWC layer ops (tensor-core mode) ----------------------------- *
 * Every entry point of this group takes `f16`: 0 = the 16-bit tensors (activations in / out, packed weights) hold
 * bfloat16 — the training path's format (gradients need the exponent range); non-zero = they hold IEEE f16 — the
 * inference path's format: 11 significand bits instead of 8 keep the per-view probabilities inside the 2e-2 bound,
 * which bf16 operands miss on this 22-layer network (2.0-2.3e-2 at the worst pixel).  tcgen05
 * kind::f16 runs both at the same rate with fp32 accumulation; conversions to f16 saturate at +-65504.  The "bf16" in
 * the names is historical: it stands for "16-bit". */

/* First layer, Cin in {1,2}: x fp32 NCHW [B,Cin,H,W] (+ optional second 1-channel
 * tensor x1 for the posterior's cat(input, segm), probabilistic_unet.py:85-90) ->
 * y bf16 NHWC [B,H,W,Cout]; w fp32 [Cout,Cin,3,3] (BN folded), bias fp32. */
int pmu_conv3x3_first_bf16(const float* x0, const float* x1, const float* w, const float* bias,
                           void* y, int B, int H, int W, int Cin, int Cout, int relu, int f16, void* stream);

/* tcgen05/TMEM implicit-GEMM convolution, TMA-fed (sm_100a only): every nn.Conv2d 3x3 + BatchNorm2d + ReLU of
 * DoubleConv / Encoder (unet_parts.py:15-20, probabilistic_unet.py:38-45), the cat of Up.forward (unet_parts.py:65-66)
 * as a two-source K loop, and nn.ConvTranspose2d k2 s2 (unet_parts.py:52).
 *  ntaps = 9: conv3x3 pad 1 over cat(x0[B,H,W,C0], x1[B,H,W,C1]) (x1 nullable);
 *             wpack bf16 [Cout][9][C0+C1] (tap = ky*3+kx); y bf16 [B,H,W,Cout].
 *  ntaps = 4: ConvTranspose2d k2 s2: x0[B,H,W,C0]; wpack bf16 [4*Cout][C0]
 *             (row = (i*2+j)*Cout + co); y bf16 [B,2H,2W,Cout]; relu must be 0.
 *  ntaps = 1: conv1x1, wpack [Cout][C0].
 *  C0, C1 multiples of 64; Cout multiple of 64; bias fp32 [Cout].
 *  f16: format of x0, x1, wpack and y (see the group comment).
 *  out_h, out_w (ntaps = 4 only, else 0): extents of the output TENSOR when it is larger than 2H x 2W — Up.forward pads
 *  the upsampled map to the skip connection's size (F.pad, unet_parts.py:58-62; an odd extent's pad row / column goes
 *  to the high side): y is then [B,out_h,out_w,Cout], the result lands at its origin and the caller zero-fills the rest. */
int pmu_conv_gemm_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack,
                       const float* bias, void* y, int B, int H, int W, int Cout, int ntaps,
                       int relu, int f16, int out_h, int out_w, void* stream);

/* conv3x3 (as above, ntaps = 9) that ALSO writes the 2x2-pooled map y_pool bf16 [B,H/2,W/2,Cout]
 * from its epilogue (pool_mode PMU_POOL_MAX: unet_parts.py:33 after DoubleConv;
 * PMU_POOL_AVG_CEIL: probabilistic_unet.py:36) — the pooled tensor never costs an extra pass
 * over HBM.  Needs even H >= 8, W >= 16.  y may be NULL when only the pooled map is consumed
 * (the prior encoder, probabilistic_unet.py:36-45). */
int pmu_conv_gemm_pool_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack,
                            const float* bias, void* y, void* y_pool, int pool_mode, int B, int H,
                            int W, int Cout, int relu, int f16, void* stream);

/* 2x2 pooling on bf16 NHWC (unet_parts.py:33 / probabilistic_unet.py:36), for shapes the fused epilogue does not take. */
int pmu_pool2_bf16(const void* x, void* y, int B, int H, int W, int C, int mode, int f16, void* stream);
/* AxisAlignedConvGaussian head on bf16 NHWC (probabilistic_unet.py:97-108): enc [B,h,w,C]; w fp32 [2L,C]; outputs fp32. */
int pmu_gauss_head_bf16(const void* enc, const float* w, const float* b, float* mu,
                        float* log_sigma, int B, int C, int h, int w_, int L, int f16, void* stream);
/* bf16 NHWC [B,H,W,C] -> fp32 NCHW [B,C,H,W]: hands unet_features back to the reference-facing API in the layout
 * ProbabilisticUnet.forward leaves it in (probabilistic_unet.py:222). */
int pmu_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int C, int f16, void* stream);

/* K3+K4 fused: fcomb over N samples with tensor-core MLP, softmax, and per-pixel
 * sum / sum-of-squares accumulation; per-sample logits never reach HBM.
 *  feat bf16 NHWC [B,H,W,64]; mu, sigma fp32 [B,L]; eps fp32 [B,N,L]
 *  (z = mu + sigma*eps, probabilistic_unet.py:233-239 rsample);
 *  w0 fp32 [64,64+L], b0[64]; wmid fp32 [nl-2][64,64], bmid; wlast [C,64], blast[C];
 *  slice_sums fp32 [B,2,C,H,W] (overwritten).  F must be 64, C <= 8, L <= 16, nl <= 6.
 *  f16: format of feat; the weights are converted to the same format inside the kernel and the hidden activations
 *  are kept in it (in tensor memory). */
int pmu_fcomb_softmax_accum_bf16(const void* feat, const float* mu, const float* sigma,
                                 const float* eps, const float* w0, const float* b0,
                                 const float* wmid, const float* bmid, const float* wlast,
                                 const float* blast, float* slice_sums, int B, int N, int L,
                                 int C, int nl, int64_t HW, int f16, void* stream);

/* ---- K4: softmax + scatter-accumulate + finalise (data plane out) -------- *
 * replaces eval.py:157 (softmax), :176-190 (cat/permute), :193 (fusion)      */

/* Unfused variant: softmax over C (eval.py:157) of logits fp32 [B,N,C,HW], summed over the N samples of the loop
 * eval.py:146-154 -> slice_sums [B,2,C,HW] = (sum p, sum p^2), overwritten. */
int pmu_softmax_accum(const float* logits, float* slice_sums, int B, int N, int C, int64_t HW,
                      void* stream);
/* S1/S2 [x][C][y][z] += slice_sums[b][0/1][C][H][W] for slices s0..s0+ns of `plane`:
 * plane 0 -> (s,r,c), plane 1 -> (r,s,c), plane 2 -> (r,c,s)  (eval.py:176,182,188). */
int pmu_scatter_accum(const float* slice_sums, int plane, int s0, int ns, const int32_t dims[3],
                      int C, float* S1, float* S2, void* stream);
/* [build-defined; closes the use_standard_axis=False TODO of utils/mri_dataset.py:60-71 for
 * the fusion side] Scatter for a NON-identity slice grid: every pixel (s, r, c) of slice_sums [ns,2,C,H,W] is added to
 * the voxel nearest to q = fma(c, v, fma(r, u, fma(s, n, o))) (affine = 12 HOST floats [o, n, u, v], the grid
 * pmu_slice_gather sampled), pixels outside the volume are dropped, cnt [x][y][z] += weight (the number of latent
 * samples behind the sums, so that mean = S1 / cnt).  fp32 atomics: several pixels can share a voxel. */
int pmu_scatter_accum_affine(const float* slice_sums, const float* affine, int s0, int ns, int H, int W,
                             const int32_t dims[3], int C, float weight, float* S1, float* S2, float* cnt,
                             void* stream);
/* pmu_fuse_finalize with a per-voxel count (the companion of pmu_scatter_accum_affine): mean = S1 / cnt,
 * var = max(S2 / cnt - mean^2, 0), entropy, argmax labels; voxels with cnt = 0 read 0 (eval.py:193 generalised). */
int pmu_fuse_finalize_counted(const float* S1, const float* S2, const float* cnt, const int32_t dims[3], int C,
                              float* mean, float* var, float* entropy, uint8_t* labels, void* stream);
/* mean = S1/count, var = max(S2/count - mean^2, 0) ([x][C][y][z]); entropy [x][y][z]
 * = sum_k -mean_k ln mean_k; labels (nullable, uint8 [x][y][z]) = argmax_k mean
 * (eval.py:52).  Any of mean/var/entropy may be NULL. */
int pmu_fuse_finalize(const float* S1, const float* S2, float count, const int32_t dims[3], int C,
                      float* mean, float* var, float* entropy, uint8_t* labels, void* stream);

/* ---- K5: reductions for the training step / evaluation ------------------- */
/* sum over b,pixels of CE(logits[B,C,HW], target[B,HW] float labels) -> out[0]
 * (probabilistic_unet.py:288,303-304).  out has TWO floats and is overwritten: out[1] > 0 reports that some label
 * was outside [0, C) — nn.CrossEntropyLoss raises "Target k is out of bounds" there, and so must the caller. */
int pmu_ce_sum(const float* logits, const float* target, int B, int C, int64_t HW, float* out,
               void* stream);
/* analytic KL(q||p) of diagonal Gaussians -> out[B] (probabilistic_unet.py:272). */
int pmu_kl_diag_gauss(const float* mu_q, const float* log_sigma_q, const float* mu_p,
                      const float* log_sigma_p, int B, int L, float* out, void* stream);
/* sums[3] = {sum(p*t), sum(p), sum(t)} (dice_loss.py:10-12); overwritten. */
int pmu_dice_sums(const float* pred, const float* target, int64_t n, float* sums, void* stream);
/* Dice of one-hot(argmax_C prob)[k] vs (truth==k) for k = 1..C-1 (eval.py:42-49):
 * prob [X][C][YZ] (avg_volume layout), truth float [X][YZ]; sums[(C-1)*3] overwritten. */
int pmu_argmax_dice_sums(const float* prob, const float* truth, int64_t X, int C, int64_t YZ,
                         float* sums, void* stream);

/* ---- training step, fp32 NCHW parity mode (train.py:85-110) ---------- *
 * What autograd does for the reference, as explicit kernels.  Data gradients of conv3x3 / conv1x1
 * reuse pmu_conv3x3_f32 / pmu_conv1x1_f32 with transposed (+ flipped) weights.  `ws` is a caller-
 * owned scratch of 2*C doubles.  Weight-gradient entry points ADD into dw (zero-fill first).   */

/* nn.BatchNorm2d in training mode (+ReLU) (unet_parts.py:16-17, probabilistic_unet.py:39-40):
 * mean/var[C] = batch statistics of y[B,C,HW] (biased variance); a = [relu](gamma*(y-mean)/sqrt(var+eps)+beta);
 * run_mean/run_var (nullable) updated with `momentum` (unbiased variance), as torch does. */
int pmu_bn_train_fwd_f32(const float* y, const float* gamma, const float* beta, float eps, int relu,
                         float momentum, float* run_mean, float* run_var, float* mean, float* var,
                         float* a, double* ws, int B, int C, int64_t HW, void* stream);
/* backward of the above (what loss.backward(), train.py:95, does for unet_parts.py:16-17): da -> dy[B,C,HW],
 * dgamma[C], dbeta[C] (overwritten; nullable). */
int pmu_bn_train_bwd_f32(const float* da, const float* y, const float* mean, const float* var,
                         const float* gamma, const float* beta, float eps, int relu, float* dy,
                         float* dgamma, float* dbeta, double* ws, int B, int C, int64_t HW, void* stream);
/* the same + dbias[C] (nullable) = sum_{b,p} dy[b,c,p], the bias gradient of the nn.Conv2d in front of the BatchNorm
 * (unet_parts.py:15-16), accumulated while dy is written; ws = 3*C doubles when dbias is given. */
int pmu_bn_train_bwd_bias_f32(const float* da, const float* y, const float* mean, const float* var,
                              const float* gamma, const float* beta, float eps, int relu, float* dy,
                              float* dgamma, float* dbeta, float* dbias, double* ws, int B, int C, int64_t HW,
                              void* stream);
/* out[c] = sum_{b,p} x[b,c,p] (bias gradients of nn.Conv2d, unet_parts.py:15,18,73; probabilistic_unet.py:137-146);
 * out[r] = sum_p x[r,p] (per-slice sums for the Fcomb layer-0 split below). */
int pmu_channel_sums_f32(const float* x, float* out, double* ws, int B, int C, int64_t HW, void* stream);
int pmu_row_sums_f32(const float* x, float* out, int64_t rows, int64_t n, void* stream);
/* dw[Cout,C0+C1,3,3] += sum_{b,h,w} dy[b,co,h,w] * cat(x0,x1)[b,ci,h+ky-1,w+kx-1]: weight gradient of the 3x3
 * nn.Conv2d layers (unet_parts.py:15,18; probabilistic_unet.py:38,43). */
int pmu_conv3x3_wgrad_f32(const float* x0, int C0, const float* x1, int C1, const float* dy, float* dw,
                          int B, int H, int W, int Cout, void* stream);
/* dw[co*ldw + ci] += sum_{b,p} dy[b,co,p] * x[b,ci,p]: weight gradient of the 1x1 nn.Conv2d layers (Fcomb,
 * probabilistic_unet.py:137-146; OutConv, unet_parts.py:73); ldw >= Cin. */
int pmu_conv1x1_wgrad_f32(const float* x, const float* dy, float* dw, int ldw, int B, int Cin, int Cout,
                          int64_t HW, void* stream);
/* MaxPool2d(2) (unet_parts.py:33) / AvgPool2d(2,2,ceil_mode) (probabilistic_unet.py:36) backward: x[B,C,H,W] (max
 * only), dy pooled -> dx[B,C,H,W]. */
int pmu_pool2_bwd_f32(const float* x, const float* dy, float* dx, int B, int C, int H, int W, int mode,
                      void* stream);
/* nn.ConvTranspose2d(k=2,s=2) backward (unet_parts.py:52): dy[B,Cout,2H,2W], w[Cin,Cout,2,2]. */
int pmu_convt2x2_dgrad_f32(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H,
                           int W, void* stream);
int pmu_convt2x2_wgrad_f32(const float* x, const float* dy, float* dw, int B, int Cin, int Cout, int H,
                           int W, void* stream);
/* dx = dy * (a > 0) (nn.ReLU backward, probabilistic_unet.py:138,143);  dst += src (gradient fan-in of the skip
 * connections, unet_model.py:49-52). */
int pmu_relu_bwd_f32(const float* a, const float* dy, float* dx, int64_t n, void* stream);
int pmu_add_f32(float* dst, const float* src, int64_t n, void* stream);
/* d(sum CE)/dlogits * scale = scale * (softmax - onehot(target))   (probabilistic_unet.py:288,303-304). */
int pmu_ce_bwd_f32(const float* logits, const float* target, float scale, float* dlogits, int B, int C,
                   int64_t HW, void* stream);
/* gradients of scale * sum_b KL(q||p) w.r.t. mu / log_sigma of both Gaussians (probabilistic_unet.py:272). */
int pmu_kl_bwd_f32(const float* mu_q, const float* ls_q, const float* mu_p, const float* ls_p, float scale,
                   float* dmu_q, float* dls_q, float* dmu_p, float* dls_p, int B, int L, void* stream);
/* backward of pmu_gauss_head_f32 (probabilistic_unet.py:97-108): denc[B,C,h,w] overwritten; dw[2L,C], db[2L] += . */
int pmu_gauss_head_bwd_f32(const float* enc, const float* w, const float* dmu, const float* dls, float* denc,
                           float* dw, float* db, int B, int C, int h, int w_, int L, void* stream);
/* Fcomb layer 0 split (probabilistic_unet.py:155-181): zb[b,co] = b0[co] + sum_l w0[co,F+l] z[b,l];
 * backward from rs[b,co] = sum_p dh0[b,co,p]: dz[B,L] overwritten; dw0[:,F:], db0 += . */
int pmu_fcomb_zbias_f32(const float* z, const float* w0, const float* b0, float* zb, int B, int F, int L,
                        void* stream);
int pmu_fcomb_zbias_bwd_f32(const float* rs, const float* z, const float* w0, float* dz, float* dw0,
                            float* db0, int B, int F, int L, void* stream);
/* y[b,co,p] = [relu](sum_ci w[co*ldw+ci] x[b,ci,p] + bias[b*bias_bstride+co])  (bias nullable): the 1x1 layers of
 * Fcomb (probabilistic_unet.py:137-146) with a per-slice bias — layer 0 after the split above — and their data gradients. */
int pmu_conv1x1_bb_f32(const float* x, const float* w, int ldw, const float* bias, int bias_bstride, float* y,
                       int B, int Cin, int Cout, int64_t HW, int relu, void* stream);

/* ---- training step, bf16 tensor-core mode ------------------------------------------------- */
/* fp32 NCHW [B,C,H,W] -> bf16 NHWC [B,H,W,C]: operand cast for the tensor-core convolutions of the training step
 * (the reference keeps NCHW fp32 throughout, train.py:85-110). */
int pmu_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int H, int W, int C, void* stream);
/* space-to-depth: x bf16 NHWC [B,2H,2W,C] -> y bf16 NHWC [B,H,W,4C], y[b,h,w,(i*2+j)*C+c] = x[b,2h+i,2w+j,c] (C % 8 == 0).
 * With it the data / weight gradients of nn.ConvTranspose2d(k=2, s=2) (unet_parts.py:52) are 1x1 GEMMs:
 * pmu_conv_gemm_bf16(ntaps = 1, K = 4*Cout) and pmu_conv_wgrad_bf16(ntaps = 1, N = 4*Cout). */
int pmu_s2d_nhwc_bf16(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* ---- training step, tensor-core mode: elementwise / reduction side on bf16 NHWC [npix = B*H*W][C] (C % 8 == 0, C/8 | 256),
 * the layout the tcgen05 GEMMs read and write — no cast between two GEMMs of the step. --------------------------------- */
/* nn.BatchNorm2d in train() mode + ReLU (unet_parts.py:16-20, probabilistic_unet.py:39-45; train.py:94): batch statistics
 * of y (one pass: per-block fp32 partials, added in fp64), running statistics updated like torch,
 * a = [relu](gamma * (y - mean) / sqrt(var + eps) + beta) as bf16.  ws: PMU_RED_MAX_BLOCKS * 2 * C floats of scratch;
 * scale_shift: 2*C floats of scratch (the folded per-channel scale / shift). */
int pmu_bn_train_fwd_nhwc_bf16(const void* y, const float* gamma, const float* beta, float eps, int relu,
                               float momentum, float* run_mean, float* run_var, float* mean, float* var,
                               void* a, float* ws, float* scale_shift, int64_t npix, int C, void* stream);
/* nn.Conv2d + the batch statistics of the train()-mode nn.BatchNorm2d behind it in ONE kernel (DoubleConv,
 * unet_parts.py:15-16,18-19; Encoder, probabilistic_unet.py:38-39,43-44; train.py:94): pmu_conv_gemm_bf16 (ntaps 9 or 1, no
 * ReLU) whose epilogue also adds, per output channel, the sum and the sum of squares of the stored (16-bit rounded)
 * outputs into stats[2*co], stats[2*co+1] (fp64; zero-filled by the caller) — the BatchNorm never re-reads y for them. */
int pmu_conv_gemm_bnstats_bf16(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                               void* y, double* stats, int B, int H, int W, int Cout, int ntaps, int f16, void* stream);
/* the rest of that BatchNorm (+ ReLU) from those sums: mean / biased variance, running statistics updated like torch,
 * a = [relu](gamma * (y - mean) / sqrt(var + eps) + beta) as bf16 (unet_parts.py:16-17).  scale_shift: 2*C floats of scratch. */
int pmu_bn_train_fwd_stats_nhwc_bf16(const void* y, const double* stats, const float* gamma, const float* beta, float eps,
                                     int relu, float momentum, float* run_mean, float* run_var, float* mean, float* var,
                                     void* a, float* scale_shift, int64_t npix, int C, void* stream);
/* its backward (loss.backward(), train.py:95): dy (bf16), dgamma, dbeta from da (bf16) and the recorded y / mean / var;
 * the ReLU mask is recomputed from y.  ws: PMU_RED_MAX_BLOCKS * 2 * C floats; coef: 4*C floats of scratch (the elementwise
 * pass is dy = coef0 * dz + coef2 * y + coef3 with the mask y * coef0 + coef1 > 0). */
int pmu_bn_train_bwd_nhwc_bf16(const void* da, const void* y, const float* mean, const float* var, const float* gamma,
                               const float* beta, float eps, int relu, void* dy, float* dgamma, float* dbeta,
                               float* ws, float* coef, int64_t npix, int C, void* stream);
/* out[s][C] = sum over the pixels of segment s of x[nseg][npix][C]: the bias gradient of nn.ConvTranspose2d
 * (unet_parts.py:52) and of the Fcomb 1x1 layers (probabilistic_unet.py:137-146) with nseg = 1; the per-slice sums behind
 * the latent part of Fcomb's first layer (probabilistic_unet.py:167-176) with nseg = B.  ws: PMU_RED_MAX_BLOCKS * C floats. */
int pmu_channel_sums_nhwc_bf16(const void* x, float* out, float* ws, int nseg, int64_t npix, int C, void* stream);
/* backward of nn.MaxPool2d(2) (unet_parts.py:33; x = the pooling input, gradient to the first maximum in row-major order
 * like torch) / nn.AvgPool2d(2, 2, ceil_mode=True) (probabilistic_unet.py:36; x may be NULL): dy [B,Ho,Wo,C] -> dx [B,H,W,C]. */
int pmu_pool2_bwd_nhwc_bf16(const void* x, const void* dy, void* dx, int B, int H, int W, int C, int mode, void* stream);
/* dst += src (bf16, n % 8 == 0): the gradient of a skip connection meets the gradient from the level below (torch.cat,
 * unet_parts.py:65). */
int pmu_add_bf16(void* dst, const void* src, int64_t n, void* stream);
/* backward of the AxisAlignedConvGaussian head (probabilistic_unet.py:97-108) on a bf16 NHWC encoder map [B,h,w,C]:
 * denc (bf16), dw [2L,C] +=, db [2L] +=  (zero-fill dw, db). */
int pmu_gauss_head_bwd_nhwc_bf16(const void* enc, const float* w, const float* dmu, const float* dls, void* denc,
                                 float* dw, float* db, int B, int C, int h, int w_, int L, void* stream);
/* ---- training step, tensor-core mode: weight layouts and the Fcomb head on bf16 NHWC ---------------------------------- */
/* nn.Conv2d weights (unet_parts.py:15,18; probabilistic_unet.py:38,43), fp32 OIHW [Cout][Cin][3][3] -> the two bf16
 * operand layouts of the tcgen05 GEMMs in one pass: wf [Cout][9][Cin] (forward, tap = ky*3+kx; nullable) and
 * wd [Cin][9][Cout] with the taps flipped (data gradient of loss.backward(), train.py:95; nullable).  Channels % 32 == 0. */
int pmu_pack_conv3x3_weights_bf16(const float* w, void* wf, void* wd, int Cout, int Cin, void* stream);
/* the same for every 3x3 layer of a network in ONE launch (the 35 conv layers of the trainer model: unet_parts.py:15,18,
 * probabilistic_unet.py:38,43): `table` is a DEVICE array of nlayers x 6 int64 {w, wf, wd, Cout, Cin, first_tile} — device
 * pointers as integers, wf / wd nullable, first_tile = running sum of (Cout/32)*(Cin/32), total_tiles = its end. */
int pmu_pack_conv3x3_weights_multi_bf16(const int64_t* table, int nlayers, int64_t total_tiles, void* stream);
/* the weight gradient pmu_conv_wgrad_bf16 produced, fp32 [Cout][9][Cin] -> the parameter's OIHW [Cout][Cin][3][3]
 * (what autograd hands to nn.Conv2d.weight.grad, train.py:95).  Cin % 64 == 0. */
int pmu_unpack_conv3x3_wgrad_f32(const float* dwp, float* dw, int Cout, int Cin, void* stream);
/* 1x1 convolution with a per-image bias on tcgen05: y[b,h,w,co] = [relu](sum_ci wpack[co][ci] x[b,h,w,ci] + bias[b][co]).
 * Fcomb's first layer (probabilistic_unet.py:167-176: the latent vector is tiled over the image and concatenated to the
 * features; its part of the 1x1 convolution is a per-slice bias W0z * z_b + b0).  Images of >= 128 pixels, channels % 64. */
int pmu_conv1x1_slicebias_bf16(const void* x, int Cin, const void* wpack, const float* bias, void* y, int B, int H,
                               int W, int Cout, int relu, int f16, void* stream);
/* Fcomb's last layer (probabilistic_unet.py:146,181: nn.Conv2d(F, n_classes, 1), no activation) from the bf16 NHWC hidden
 * map h [B*HW][F] to fp32 NCHW logits [B][C][HW] for the cross entropy (probabilistic_unet.py:294-299).  C <= 8. */
int pmu_fcomb_last_fwd_bf16(const void* h, const float* w, const float* bias, float* logits, int B, int64_t HW, int F,
                            int C, void* stream);
/* its backward in loss.backward() (train.py:95): dh (bf16 NHWC) = (h > 0) * W^T dlogits — the ReLU in front folded in —
 * and dw [C][F] = sum_p dlogits[k] h[c] (written, not accumulated).  ws: PMU_RED_MAX_BLOCKS * C * F floats.  C <= 4. */
int pmu_fcomb_last_bwd_bf16(const void* h, const float* dlogits, const float* w, void* dh, float* dw, float* ws, int B,
                            int64_t HW, int F, int C, void* stream);
/* d = (h > 0) ? d : 0 in place (bf16, n % 8 == 0): backward of the nn.ReLU between Fcomb's 1x1 layers
 * (probabilistic_unet.py:141-144). */
int pmu_relu_mask_bf16(void* d, const void* h, int64_t n, void* stream);
/* weight gradient of the first convolution of the U-Net / prior (Cin = 1: x1 NULL) / posterior (Cin = 2: image x0 and
 * mask x1, the torch.cat of probabilistic_unet.py:88-92 never materialised) in the tensor-core training step:
 * dw fp32 OIHW [Cout][Cin][3][3] (written) from the fp32 NCHW inputs and the bf16 NHWC gradient dy [B,H,W,Cout]
 * (unet_parts.py:15 / probabilistic_unet.py:38 in loss.backward(), train.py:95).  ws: PMU_RED_MAX_BLOCKS * 9 * Cout floats. */
int pmu_conv3x3_wgrad_smallcin_bf16(const float* x0, const float* x1, const void* dy, float* dw, float* ws, int B, int H,
                                    int W, int Cout, void* stream);
/* tcgen05 weight gradient of conv3x3 pad 1 (ntaps = 9) / conv1x1 (ntaps = 1):
 * dw fp32 [Cout][ntaps][C0+C1] (+)= sum_{b,h,w} dy[b,h,w,co] * cat(x0,x1)[b,h+ky-1,w+kx-1,ci]   (tap = ky*3+kx)
 * x0 bf16 [B,H,W,C0], x1 (nullable) bf16 [B,H,W,C1], dy bf16 [B,H,W,Cout]; channels multiples of 64.
 * overwrite = 0: dw += (zero-fill it for a plain gradient); the reduction over pixels is split across CTAs and partials are
 * added with fp32 atomics.  overwrite = 1: dw is written (no zero-fill by the caller): tiles that are not worth splitting —
 * the layers with many (tap, channel) tiles, i.e. the large weights — leave with plain stores, split ones after a memset
 * by the library.
 * Weight gradient of the nn.Conv2d layers of unet_parts.py:15,18 / probabilistic_unet.py:38,43 in loss.backward() (train.py:95). */
int pmu_conv_wgrad_bf16(const void* x0, int C0, const void* x1, int C1, const void* dy, float* dw, int B,
                        int H, int W, int Cout, int ntaps, int overwrite, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMU_B200_H_ */
