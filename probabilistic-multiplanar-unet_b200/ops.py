"""Tensor-level wrappers over the C-ABI (include/pmu_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic step runs in
libpmu_b200.so.  Each wrapper passes raw device pointers + the current CUDA stream.  There is
no fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int32, c_void_p
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

INTERP = {"exact": 0, "nearest": 1, "trilinear": 2}
POOL_MAX, POOL_AVG_CEIL = 0, 1


# ---- instrumentation (bench.py): how many of OUR kernels were launched, and optional per-call
# CUDA-event timing on the launching stream -------------------------------------------------
LAUNCHES = 0
PROFILE = None      # set to a list to record (name, meta, start_event, end_event) per C-ABI call
_META = None


def _launch(lib, name, args):
    global LAUNCHES, _META
    meta, _META = _META, None
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.append((name, meta, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    _lib.check(rc, name)
    LAUNCHES += 1


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _prep(*tensors: Optional[torch.Tensor]):
    """Validate tensors, select the device inside the library, return (lib, stream)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pmu_b200 ops need CUDA tensors (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("pmu_b200 ops need contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("pmu_b200 ops: tensors on different devices")
    lib = _lib.load()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    # the library has its own (statically linked) CUDA runtime with a per-thread current device; binding the device's
    # launch context selects it and gives the launch its cached TMA descriptors / kernel attributes
    _lib.bind_device(lib, idx)
    return lib, c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dims(d: Sequence[int]):
    return (c_int32 * 3)(int(d[0]), int(d[1]), int(d[2]))


def _f32(t, name):
    if t is not None and t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")


def _bf16(t, name):
    if t is not None and t.dtype != torch.bfloat16:
        raise RuntimeError(f"{name} must be bfloat16, got {t.dtype}")


def _h16(*named) -> int:
    """The 16-bit tensors of a tensor-core op — (tensor, name) pairs, None tensors skipped — must share ONE format:
    bfloat16 (training path) or float16 (inference path).  Returns the `f16` flag of the C-ABI."""
    dt = None
    for t, name in named:
        if t is None:
            continue
        if t.dtype not in (torch.bfloat16, torch.float16):
            raise RuntimeError(f"{name} must be bfloat16 or float16, got {t.dtype}")
        if dt is not None and t.dtype != dt:
            raise RuntimeError(f"{name} is {t.dtype} but the other 16-bit operands are {dt}: one format per call")
        dt = t.dtype
    return int(dt == torch.float16)


# ----------------------------------------------------------------------------- K1
def plane_max(vol: torch.Tensor) -> torch.Tensor:
    """Per-slice maxima of all three planes, one pass: [d0 + d1 + d2] fp32."""
    _f32(vol, "vol")
    lib, st = _prep(vol)
    d = vol.shape
    out = torch.empty(d[0] + d[1] + d[2], dtype=torch.float32, device=vol.device)
    _launch(lib, "pmu_fill_f32", (_p(out), float("-inf"), out.numel(), st,))
    _launch(lib, "pmu_plane_max", (_p(vol), _dims(d), _p(out), st,))
    return out


def slice_gather(vol: torch.Tensor, plane: int, s0: int, ns: int, *, interp: str = "exact",
                 affine: Optional[Sequence[float]] = None, hw: Optional[Tuple[int, int]] = None,
                 slice_max_in: Optional[torch.Tensor] = None, want_max: bool = False,
                 out: Optional[torch.Tensor] = None):
    """out[ns,1,H,W] fp32 (+ per-slice raw max [ns] when want_max)."""
    _f32(vol, "vol"); _f32(slice_max_in, "slice_max_in")
    d = vol.shape
    if hw is None:
        hw = (d[1] if plane == 0 else d[0], d[1] if plane == 2 else d[2])
    H, W = hw
    if out is None:
        out = torch.empty(ns, 1, H, W, dtype=torch.float32, device=vol.device)
    lib, st = _prep(vol, slice_max_in, out)
    mx = None
    if want_max:
        mx = torch.empty(max(ns, 1), dtype=torch.float32, device=vol.device)
        _launch(lib, "pmu_fill_f32", (_p(mx), float("-inf"), mx.numel(), st,))
    aff = None
    if affine is not None:
        aff = (c_float * 12)(*[float(a) for a in affine])
    _launch(lib, "pmu_slice_gather", (_p(vol), _dims(d), int(plane), int(s0), int(ns), INTERP[interp], aff, int(H),
                                    int(W), _p(slice_max_in), _p(mx), _p(out), st,))
    return (out, mx[:ns]) if want_max else out


def slice_normalize_(slices: torch.Tensor, slice_max: torch.Tensor) -> torch.Tensor:
    _f32(slices, "slices"); _f32(slice_max, "slice_max")
    lib, st = _prep(slices, slice_max)
    ns = slices.shape[0]
    _launch(lib, "pmu_slice_normalize", (_p(slices), _p(slice_max), ns, slices.numel() // max(ns, 1), st,))
    return slices


# ----------------------------------------------------------------------------- fp32 layers
def conv3x3_f32(x0, w, bias, relu=True, x1=None, out=None):
    _f32(x0, "x0"); _f32(x1, "x1"); _f32(w, "w"); _f32(bias, "bias")
    B, C0, H, W = x0.shape
    C1 = 0 if x1 is None else x1.shape[1]
    Cout = w.shape[0]
    assert w.shape[1] == C0 + C1, (w.shape, C0, C1)
    if out is None:
        out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x0.device)
    lib, st = _prep(x0, x1, w, bias, out)
    _launch(lib, "pmu_conv3x3_f32", (_p(x0), C0, _p(x1), C1, _p(w), _p(bias), _p(out), B, H, W, Cout, int(relu), st,))
    return out


def conv1x1_f32(x, w, bias, relu=False):
    _f32(x, "x"); _f32(w, "w")
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w, bias, out)
    _launch(lib, "pmu_conv1x1_f32", (_p(x), _p(w), _p(bias), _p(out), B, Cin, Cout, H * W, int(relu), st,))
    return out


def convt2x2_f32(x, w, bias, out_hw=None):
    _f32(x, "x"); _f32(w, "w")
    B, Cin, H, W = x.shape
    Cout = w.shape[1]
    Ho, Wo = (2 * H, 2 * W) if out_hw is None else out_hw
    dy, dx = Ho - 2 * H, Wo - 2 * W
    out = torch.empty(B, Cout, Ho, Wo, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w, bias, out)
    _launch(lib, "pmu_convt2x2_f32", (_p(x), _p(w), _p(bias), _p(out), B, Cin, Cout, H, W, Ho, Wo, dy // 2, dx // 2, st,))
    return out


def pool2_f32(x, mode):
    _f32(x, "x")
    B, C, H, W = x.shape
    Ho, Wo = (H // 2, W // 2) if mode == POOL_MAX else ((H + 1) // 2, (W + 1) // 2)
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _launch(lib, "pmu_pool2_f32", (_p(x), _p(out), B, C, H, W, mode, st,))
    return out


def gauss_head_f32(enc, w, b, L):
    _f32(enc, "enc")
    B, C, h, w_ = enc.shape
    mu = torch.empty(B, L, dtype=torch.float32, device=enc.device)
    ls = torch.empty_like(mu)
    lib, st = _prep(enc, w, b, mu, ls)
    _launch(lib, "pmu_gauss_head_f32", (_p(enc), _p(w), _p(b), _p(mu), _p(ls), B, C, h, w_, L, st,))
    return mu, ls


def fcomb_f32(feat, z, fw, want_logits=True, want_sums=False, sums_out=None):
    """feat [B,F,H,W] fp32, z [B,N,L] -> logits [B,N,C,H,W] and/or slice_sums [B,2,C,H,W] (into sums_out when given)."""
    _f32(feat, "feat"); _f32(z, "z"); _f32(sums_out, "sums_out")
    B, F_, H, W = feat.shape
    N, L = z.shape[1], z.shape[2]
    C, nl = fw["wlast"].shape[0], fw["nl"]
    logits = torch.empty(B, N, C, H, W, dtype=torch.float32, device=feat.device) if want_logits else None
    sums = None
    if want_sums:
        if sums_out is not None and tuple(sums_out.shape) != (B, 2, C, H, W):
            raise RuntimeError(f"fcomb_f32: sums_out must be {(B, 2, C, H, W)}, got {tuple(sums_out.shape)}")
        sums = sums_out if sums_out is not None else torch.empty(B, 2, C, H, W, dtype=torch.float32, device=feat.device)
    lib, st = _prep(feat, z, logits, sums, fw["w0"], fw["b0"], fw["wmid"], fw["bmid"], fw["wlast"], fw["blast"])
    _launch(lib, "pmu_fcomb_f32", (_p(feat), _p(z), _p(fw["w0"]), _p(fw["b0"]), _p(fw["wmid"]), _p(fw["bmid"]),
                                 _p(fw["wlast"]), _p(fw["blast"]), _p(logits), _p(sums), B, N, F_, L, C, nl, H * W, st,))
    return logits, sums


# ----------------------------------------------------------------------------- bf16 layers
def conv3x3_first_bf16(x0, w, bias, relu=True, x1=None, out_dtype=torch.bfloat16):
    """fp32 NCHW input -> 16-bit NHWC output in `out_dtype` (bfloat16 or float16)."""
    _f32(x0, "x0"); _f32(x1, "x1")
    B, _, H, W = x0.shape
    Cin = 1 if x1 is None else 2
    Cout = w.shape[0]
    out = torch.empty(B, H, W, Cout, dtype=out_dtype, device=x0.device)
    f16 = _h16((out, "out_dtype"))
    lib, st = _prep(x0, x1, w, bias, out)
    _launch(lib, "pmu_conv3x3_first_bf16", (_p(x0), _p(x1), _p(w), _p(bias), _p(out), B, H, W, Cin, Cout, int(relu), f16, st,))
    return out


def conv_gemm_bf16(x0, wpack, bias, Cout, ntaps, relu, x1=None, out_hw=None):
    """tcgen05 implicit-GEMM conv: x NHWC 16-bit -> y NHWC in the same format (2H x 2W for ntaps=4; out_hw = the skip
    connection's size when Up.forward pads the upsampled map, unet_parts.py:58-62: zero rows / columns at the high side)."""
    _f32(bias, "bias")
    wf = _h16((x0, "x0"), (x1, "x1"), (wpack, "wpack"))
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    oh, ow = (2 * H, 2 * W) if ntaps == 4 else (H, W)
    padded = ntaps == 4 and out_hw is not None and tuple(out_hw) != (oh, ow)
    if padded:
        if out_hw[0] < oh or out_hw[1] < ow:
            raise RuntimeError(f"conv_gemm_bf16: out_hw {tuple(out_hw)} smaller than the upsampled map {(oh, ow)}")
        out = torch.zeros(B, out_hw[0], out_hw[1], Cout, dtype=x0.dtype, device=x0.device)
    else:
        out = torch.empty(B, oh, ow, Cout, dtype=x0.dtype, device=x0.device)
    lib, st = _prep(x0, x1, wpack, bias, out)
    global _META
    ktot = (9 if ntaps == 9 else 1) * (C0 + C1)
    ntot = (4 if ntaps == 4 else 1) * Cout
    _META = {"flops": 2.0 * B * H * W * ntot * ktot,
             "bytes": 2.0 * (B * H * W * (C0 + C1) + out.numel() + ntot * ktot)}
    _launch(lib, "pmu_conv_gemm_bf16", (_p(x0), C0, _p(x1), C1, _p(wpack), _p(bias), _p(out), B, H, W, Cout, ntaps,
                                      int(relu), wf, int(out.shape[1]) if padded else 0, int(out.shape[2]) if padded else 0, st,))
    return out


def conv_gemm_pool_bf16(x0, wpack, bias, Cout, relu, pool_mode, x1=None, want_full=True):
    """conv3x3 + fused 2x2 pooling epilogue: returns (y [B,H,W,Cout] or None, y_pool [B,H/2,W/2,Cout])."""
    _f32(bias, "bias")
    wf = _h16((x0, "x0"), (x1, "x1"), (wpack, "wpack"))
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    out = torch.empty(B, H, W, Cout, dtype=x0.dtype, device=x0.device) if want_full else None
    outp = torch.empty(B, H // 2, W // 2, Cout, dtype=x0.dtype, device=x0.device)
    lib, st = _prep(x0, x1, wpack, bias, out, outp)
    global _META
    ktot = 9 * (C0 + C1)
    _META = {"flops": 2.0 * B * H * W * Cout * ktot,
             "bytes": 2.0 * (B * H * W * (C0 + C1) + (out.numel() if want_full else 0) + outp.numel() + Cout * ktot)}
    _launch(lib, "pmu_conv_gemm_pool_bf16", (_p(x0), C0, _p(x1), C1, _p(wpack), _p(bias), _p(out), _p(outp), int(pool_mode),
                                             B, H, W, Cout, int(relu), wf, st,))
    return out, outp


def fused_pool_ok(H, W):
    return H % 2 == 0 and W % 2 == 0 and H >= 8 and W >= 16


def pool2_bf16(x, mode):
    f16 = _h16((x, "x"))
    B, H, W, C = x.shape
    Ho, Wo = (H // 2, W // 2) if mode == POOL_MAX else ((H + 1) // 2, (W + 1) // 2)
    out = torch.empty(B, Ho, Wo, C, dtype=x.dtype, device=x.device)
    lib, st = _prep(x, out)
    _launch(lib, "pmu_pool2_bf16", (_p(x), _p(out), B, H, W, C, mode, f16, st,))
    return out


def gauss_head_bf16(enc, w, b, L):
    f16 = _h16((enc, "enc"))
    B, h, w_, C = enc.shape
    mu = torch.empty(B, L, dtype=torch.float32, device=enc.device)
    ls = torch.empty_like(mu)
    lib, st = _prep(enc, w, b, mu, ls)
    _launch(lib, "pmu_gauss_head_bf16", (_p(enc), _p(w), _p(b), _p(mu), _p(ls), B, C, h, w_, L, f16, st,))
    return mu, ls


def nhwc_bf16_to_nchw_f32(x):
    f16 = _h16((x, "x"))
    B, H, W, C = x.shape
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _launch(lib, "pmu_nhwc_bf16_to_nchw_f32", (_p(x), _p(out), B, H, W, C, f16, st,))
    return out


def fcomb_softmax_accum_bf16(feat, mu, sigma, eps, fw, out=None):
    """feat NHWC 16-bit [B,H,W,64]; mu/sigma [B,L]; eps [B,N,L] -> slice_sums [B,2,C,H,W] fp32."""
    f16 = _h16((feat, "feat")); _f32(mu, "mu"); _f32(sigma, "sigma"); _f32(eps, "eps")
    B, H, W, F_ = feat.shape
    if F_ != 64:
        raise RuntimeError(f"fcomb_softmax_accum_bf16 needs 64 feature channels, got {F_}")
    N, L = eps.shape[1], eps.shape[2]
    C, nl = fw["wlast"].shape[0], fw["nl"]
    if out is None:
        out = torch.empty(B, 2, C, H, W, dtype=torch.float32, device=feat.device)
    lib, st = _prep(feat, mu, sigma, eps, out)
    global _META
    nmid = nl - 2
    # algorithmic work: shared layer-0 GEMM once per pixel, hidden layers + head per sample; and the
    # accumulator read-back that bounds the kernel: G once, one 64-column fp32 tile per hidden layer
    # and the 8 head columns per sample
    _META = {"flops": 2.0 * B * H * W * (64 * 64 + N * (nmid * 64 * 64 + 64 * C)),
             "bytes": 2.0 * B * H * W * 64 + 4.0 * out.numel(),
             "tmem_read_bytes": 4.0 * B * H * W * (64 + N * (nmid * 64 + 8))}
    _launch(lib, "pmu_fcomb_softmax_accum_bf16", (_p(feat), _p(mu), _p(sigma), _p(eps), _p(fw["w0"]), _p(fw["b0"]),
                                                _p(fw["wmid"]), _p(fw["bmid"]), _p(fw["wlast"]), _p(fw["blast"]),
                                                _p(out), B, N, L, C, nl, H * W, f16, st,))
    return out


# ----------------------------------------------------------------------------- K4
def softmax_accum(logits):
    _f32(logits, "logits")
    B, N, C, H, W = logits.shape
    out = torch.empty(B, 2, C, H, W, dtype=torch.float32, device=logits.device)
    lib, st = _prep(logits, out)
    _launch(lib, "pmu_softmax_accum", (_p(logits), _p(out), B, N, C, H * W, st,))
    return out


def scatter_accum_(slice_sums, plane, s0, dims, S1, S2):
    _f32(slice_sums, "slice_sums"); _f32(S1, "S1"); _f32(S2, "S2")
    ns, _, C = slice_sums.shape[:3]
    lib, st = _prep(slice_sums, S1, S2)
    _launch(lib, "pmu_scatter_accum", (_p(slice_sums), int(plane), int(s0), int(ns), _dims(dims), int(C), _p(S1), _p(S2), st,))


def scatter_accum_affine_(slice_sums, affine, s0, dims, S1, S2, cnt, weight):
    """Non-identity slice grid (App. A step 6): nearest-voxel scatter of slice_sums [ns,2,C,H,W] with a per-voxel count
    (cnt [x,y,z] += weight)."""
    _f32(slice_sums, "slice_sums"); _f32(S1, "S1"); _f32(S2, "S2"); _f32(cnt, "cnt")
    ns, _, C, H, W = slice_sums.shape
    lib, st = _prep(slice_sums, S1, S2, cnt)
    aff = (c_float * 12)(*[float(a) for a in affine])
    _launch(lib, "pmu_scatter_accum_affine", (_p(slice_sums), aff, int(s0), int(ns), int(H), int(W), _dims(dims), int(C),
                                             float(weight), _p(S1), _p(S2), _p(cnt), st,))


def fuse_finalize_counted(S1, S2, cnt, want_var=True, want_entropy=True, want_labels=False):
    """mean / var / entropy / labels with a per-voxel count (voxels nothing landed on read 0)."""
    _f32(S1, "S1"); _f32(S2, "S2"); _f32(cnt, "cnt")
    X, C, Y, Z = S1.shape
    mean = torch.empty_like(S1)
    var = torch.empty_like(S1) if want_var else None
    ent = torch.empty(X, Y, Z, dtype=torch.float32, device=S1.device) if want_entropy else None
    lab = torch.empty(X, Y, Z, dtype=torch.uint8, device=S1.device) if want_labels else None
    lib, st = _prep(S1, S2, cnt, mean, var, ent, lab)
    _launch(lib, "pmu_fuse_finalize_counted", (_p(S1), _p(S2), _p(cnt), _dims((X, Y, Z)), C, _p(mean), _p(var), _p(ent), _p(lab), st,))
    return mean, var, ent, lab


def fuse_finalize(S1, S2, count, want_var=True, want_entropy=True, want_labels=False, out=None):
    """out = (mean, var, entropy, labels) writes into caller tensors (x-slabs of the full outputs)."""
    _f32(S1, "S1"); _f32(S2, "S2")
    X, C, Y, Z = S1.shape
    if out is not None:
        mean, var, ent, lab = out
    else:
        mean = torch.empty_like(S1)
        var = torch.empty_like(S1) if want_var else None
        ent = torch.empty(X, Y, Z, dtype=torch.float32, device=S1.device) if want_entropy else None
        lab = torch.empty(X, Y, Z, dtype=torch.uint8, device=S1.device) if want_labels else None
    lib, st = _prep(S1, S2, mean, var, ent, lab)
    _launch(lib, "pmu_fuse_finalize", (_p(S1), _p(S2), float(count), _dims((X, Y, Z)), C, _p(mean), _p(var), _p(ent),
                                     _p(lab), st,))
    return mean, var, ent, lab


# ----------------------------------------------------------------------------- K5
CE_CHECK_LABELS = True      # False while a CUDA graph is being captured (the read-back below is a host synchronisation)
CE_LAST_FLAG = None         # the device-side out-of-range flag of the last call made with the check off


def ce_sum(logits, target):
    """sum over batch and pixels of CE(logits [B,C,H,W], target float labels [B,1,H,W] or [B,H,W])."""
    _f32(logits, "logits"); _f32(target, "target")
    B, C = logits.shape[:2]
    HW = logits.numel() // (B * C)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    lib, st = _prep(logits, target, out)
    _launch(lib, "pmu_ce_sum", (_p(logits), _p(target), B, C, HW, _p(out), st,))
    # nn.CrossEntropyLoss raises on a target outside [0, C) (probabilistic_unet.py:288,303): so do we — one 4-byte
    # read-back per loss evaluation instead of training on garbage labels (255, a 4-class map on a 3-class net, ...)
    global CE_LAST_FLAG
    if not CE_CHECK_LABELS:
        CE_LAST_FLAG = out[1]
        return out[0]
    if float(out[1]) > 0:
        raise IndexError(f"Target out of bounds: labels must lie in [0, {C}) (mask values outside the class range)")
    return out[0]


def kl_diag_gauss(mu_q, ls_q, mu_p, ls_p):
    B, L = mu_q.shape
    out = torch.empty(B, dtype=torch.float32, device=mu_q.device)
    lib, st = _prep(mu_q, ls_q, mu_p, ls_p, out)
    _launch(lib, "pmu_kl_diag_gauss", (_p(mu_q), _p(ls_q), _p(mu_p), _p(ls_p), B, L, _p(out), st,))
    return out


def dice_sums(pred, target):
    _f32(pred, "pred"); _f32(target, "target")
    if pred.numel() != target.numel():
        raise RuntimeError("dice_sums: pred and target must have the same number of elements")
    out = torch.empty(3, dtype=torch.float32, device=pred.device)
    lib, st = _prep(pred, target, out)
    _launch(lib, "pmu_dice_sums", (_p(pred), _p(target), pred.numel(), _p(out), st,))
    return out


def argmax_dice_sums(prob, truth):
    """prob [X,C,Y,Z], truth float [X,Y,Z] -> sums [(C-1),3] = (inter, pred, truth) for k=1..C-1."""
    _f32(prob, "prob"); _f32(truth, "truth")
    X, C = prob.shape[:2]
    YZ = prob.numel() // (X * C)
    out = torch.empty((C - 1) * 3, dtype=torch.float32, device=prob.device)
    lib, st = _prep(prob, truth, out)
    _launch(lib, "pmu_argmax_dice_sums", (_p(prob), _p(truth), X, C, YZ, _p(out), st,))
    return out.view(C - 1, 3)


# ----------------------------------------------------------------------------- training step (fp32)
def _ws(C, dev):
    return torch.empty(2 * C, dtype=torch.float64, device=dev)


RED_MAX_BLOCKS = 592      # PMU_RED_MAX_BLOCKS of include/pmu_b200.h


def _red_ws(row_floats, dev):
    """Workspace of a two-level per-channel reduction: one fp32 row of partials per block of the first pass."""
    return torch.empty(RED_MAX_BLOCKS * row_floats, dtype=torch.float32, device=dev)


def bn_train_fwd_f32(y, gamma, beta, eps, relu, momentum=0.1, run_mean=None, run_var=None):
    """Train-mode BatchNorm2d (+ReLU): returns (a, mean, var); running stats updated in place."""
    _f32(y, "y")
    B, C, H, W = y.shape
    mean = torch.empty(C, dtype=torch.float32, device=y.device)
    var = torch.empty_like(mean)
    a = torch.empty_like(y)
    ws = _ws(C, y.device)
    lib, st = _prep(y, gamma, beta, run_mean, run_var, mean, var, a, ws)
    _launch(lib, "pmu_bn_train_fwd_f32", (_p(y), _p(gamma), _p(beta), float(eps), int(relu), float(momentum),
                                        _p(run_mean), _p(run_var), _p(mean), _p(var), _p(a), _p(ws), B, C, H * W, st,))
    return a, mean, var


def bn_train_bwd_f32(da, y, mean, var, gamma, beta, eps, relu, want_dbias=False):
    """-> (dy, dgamma, dbeta) [+ dbias = channel sums of dy, the bias gradient of the conv in front, fused]."""
    B, C, H, W = y.shape
    dy = torch.empty_like(y)
    dg = torch.empty(C, dtype=torch.float32, device=y.device)
    db = torch.empty_like(dg)
    dbias = torch.empty_like(dg) if want_dbias else None
    ws = torch.empty(3 * C, dtype=torch.float64, device=y.device)
    lib, st = _prep(da, y, mean, var, gamma, beta, dy, dg, db, ws)
    _launch(lib, "pmu_bn_train_bwd_bias_f32", (_p(da), _p(y), _p(mean), _p(var), _p(gamma), _p(beta), float(eps), int(relu),
                                             _p(dy), _p(dg), _p(db), _p(dbias), _p(ws), B, C, H * W, st,))
    return (dy, dg, db, dbias) if want_dbias else (dy, dg, db)


def channel_sums_f32(x):
    """[B,C,H,W] -> [C] sums over batch and pixels."""
    B, C = x.shape[0], x.shape[1]
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = _ws(C, x.device)
    lib, st = _prep(x, out, ws)
    _launch(lib, "pmu_channel_sums_f32", (_p(x), _p(out), _p(ws), B, C, x.numel() // (B * C), st,))
    return out


def row_sums_f32(x, rows):
    out = torch.empty(rows, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _launch(lib, "pmu_row_sums_f32", (_p(x), _p(out), rows, x.numel() // rows, st,))
    return out


def conv3x3_wgrad_f32(x0, dy, dw, x1=None):
    """dw[Cout,C0+C1,3,3] += weight gradient."""
    B, C0, H, W = x0.shape
    C1 = 0 if x1 is None else x1.shape[1]
    Cout = dy.shape[1]
    assert tuple(dw.shape) == (Cout, C0 + C1, 3, 3), dw.shape
    lib, st = _prep(x0, x1, dy, dw)
    _launch(lib, "pmu_conv3x3_wgrad_f32", (_p(x0), C0, _p(x1), C1, _p(dy), _p(dw), B, H, W, Cout, st,))


def conv1x1_wgrad_f32(x, dy, dw, ldw=None):
    B, Cin = x.shape[0], x.shape[1]
    Cout = dy.shape[1]
    lib, st = _prep(x, dy, dw)
    _launch(lib, "pmu_conv1x1_wgrad_f32", (_p(x), _p(dy), _p(dw), int(ldw or Cin), B, Cin, Cout,
                                         x.numel() // (B * Cin), st,))


def pool2_bwd_f32(x, dy, mode):
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    lib, st = _prep(x, dy, dx)
    _launch(lib, "pmu_pool2_bwd_f32", (_p(x), _p(dy), _p(dx), B, C, H, W, mode, st,))
    return dx


def convt2x2_dgrad_f32(dy, w):
    B, Cout, H2, W2 = dy.shape
    Cin = w.shape[0]
    dx = torch.empty(B, Cin, H2 // 2, W2 // 2, dtype=torch.float32, device=dy.device)
    lib, st = _prep(dy, w, dx)
    _launch(lib, "pmu_convt2x2_dgrad_f32", (_p(dy), _p(w), _p(dx), B, Cin, Cout, H2 // 2, W2 // 2, st,))
    return dx


def convt2x2_wgrad_f32(x, dy, dw):
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    lib, st = _prep(x, dy, dw)
    _launch(lib, "pmu_convt2x2_wgrad_f32", (_p(x), _p(dy), _p(dw), B, Cin, Cout, H, W, st,))


def relu_bwd_f32(a, dy):
    dx = torch.empty_like(dy)
    lib, st = _prep(a, dy, dx)
    _launch(lib, "pmu_relu_bwd_f32", (_p(a), _p(dy), _p(dx), dy.numel(), st,))
    return dx


def add_f32_(dst, src):
    assert dst.shape == src.shape
    lib, st = _prep(dst, src)
    _launch(lib, "pmu_add_f32", (_p(dst), _p(src), dst.numel(), st,))
    return dst


def ce_bwd_f32(logits, target, scale):
    B, C = logits.shape[0], logits.shape[1]
    dl = torch.empty_like(logits)
    lib, st = _prep(logits, target, dl)
    _launch(lib, "pmu_ce_bwd_f32", (_p(logits), _p(target), float(scale), _p(dl), B, C, logits.numel() // (B * C), st,))
    return dl


def kl_bwd_f32(mu_q, ls_q, mu_p, ls_p, scale):
    outs = [torch.empty_like(mu_q) for _ in range(4)]
    lib, st = _prep(mu_q, ls_q, mu_p, ls_p, *outs)
    _launch(lib, "pmu_kl_bwd_f32", (_p(mu_q), _p(ls_q), _p(mu_p), _p(ls_p), float(scale), *[_p(o) for o in outs],
                                  mu_q.shape[0], mu_q.shape[1], st,))
    return outs


def gauss_head_bwd_f32(enc, w, dmu, dls, dw, db):
    B, C, h, w_ = enc.shape
    denc = torch.empty_like(enc)
    lib, st = _prep(enc, w, dmu, dls, denc, dw, db)
    _launch(lib, "pmu_gauss_head_bwd_f32", (_p(enc), _p(w), _p(dmu), _p(dls), _p(denc), _p(dw), _p(db), B, C, h, w_,
                                          dmu.shape[1], st,))
    return denc


def fcomb_zbias_f32(z, w0, b0):
    B, L = z.shape
    F_ = w0.shape[0]
    zb = torch.empty(B, F_, dtype=torch.float32, device=z.device)
    lib, st = _prep(z, w0, b0, zb)
    _launch(lib, "pmu_fcomb_zbias_f32", (_p(z), _p(w0), _p(b0), _p(zb), B, F_, L, st,))
    return zb


def fcomb_zbias_bwd_f32(rs, z, w0, dw0, db0):
    B, L = z.shape
    F_ = w0.shape[0]
    dz = torch.empty_like(z)
    lib, st = _prep(rs, z, w0, dz, dw0, db0)
    _launch(lib, "pmu_fcomb_zbias_bwd_f32", (_p(rs), _p(z), _p(w0), _p(dz), _p(dw0), _p(db0), B, F_, L, st,))
    return dz


def conv1x1_bb_f32(x, w, ldw, bias, bias_bstride, Cin, Cout, relu):
    """y[b,co] = [relu](sum_ci w[co*ldw+ci] x[b,ci] + bias[b*bias_bstride+co])."""
    B, H, W = x.shape[0], x.shape[2], x.shape[3]
    y = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w, bias, y)
    _launch(lib, "pmu_conv1x1_bb_f32", (_p(x), _p(w), int(ldw), _p(bias), int(bias_bstride), _p(y), B, Cin, Cout, H * W,
                                      int(relu), st,))
    return y


# ----------------------------------------------------------------------------- training step (bf16 tensor-core mode)
def nchw_f32_to_nhwc_bf16(x):
    _f32(x, "x")
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x.device)
    lib, st = _prep(x, y)
    _launch(lib, "pmu_nchw_f32_to_nhwc_bf16", (_p(x), _p(y), B, H, W, C, st,))
    return y


def s2d_nhwc_bf16(x):
    """bf16 NHWC [B,2H,2W,C] -> [B,H,W,4C] with channel order (i, j, c): y[b,h,w,(i*2+j)*C+c] = x[b,2h+i,2w+j,c]."""
    _bf16(x, "x")
    B, H2, W2, C = x.shape
    if H2 % 2 or W2 % 2:
        raise RuntimeError("s2d_nhwc_bf16 needs even H and W")
    y = torch.empty(B, H2 // 2, W2 // 2, 4 * C, dtype=torch.bfloat16, device=x.device)
    lib, st = _prep(x, y)
    _launch(lib, "pmu_s2d_nhwc_bf16", (_p(x), _p(y), B, H2 // 2, W2 // 2, C, st,))
    return y


def conv_wgrad_bf16(x0, dy, dw, x1=None, ntaps=9, overwrite=False):
    """dw fp32 [Cout, ntaps, C0+C1] += tcgen05 weight gradient (overwrite: dw = ..., no zero-fill needed); x0/x1/dy bf16 NHWC."""
    _bf16(x0, "x0"); _bf16(x1, "x1"); _bf16(dy, "dy"); _f32(dw, "dw")
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    Cout = dy.shape[3]
    assert tuple(dw.shape) == (Cout, ntaps, C0 + C1), dw.shape
    lib, st = _prep(x0, x1, dy, dw)
    global _META
    _META = {"flops": 2.0 * B * H * W * Cout * ntaps * (C0 + C1)}
    _launch(lib, "pmu_conv_wgrad_bf16", (_p(x0), C0, _p(x1), C1, _p(dy), _p(dw), B, H, W, Cout, int(ntaps), int(overwrite), st,))


# ----------------------------------------------------------------------------- training step on bf16 NHWC activations
def bn_train_fwd_nhwc_bf16(y, gamma, beta, eps, relu, momentum=0.1, run_mean=None, run_var=None):
    """Train-mode BatchNorm2d (+ReLU) on y bf16 [B,H,W,C]: returns (a bf16, mean, var); running stats updated in place."""
    _bf16(y, "y")
    C = y.shape[-1]
    npix = y.numel() // C
    mean = torch.empty(C, dtype=torch.float32, device=y.device)
    var = torch.empty_like(mean)
    a = torch.empty_like(y)
    ws = _red_ws(2 * C, y.device)
    ss = torch.empty(2 * C, dtype=torch.float32, device=y.device)
    lib, st = _prep(y, gamma, beta, run_mean, run_var, mean, var, a, ws, ss)
    _launch(lib, "pmu_bn_train_fwd_nhwc_bf16", (_p(y), _p(gamma), _p(beta), float(eps), int(relu), float(momentum), _p(run_mean),
                                              _p(run_var), _p(mean), _p(var), _p(a), _p(ws), _p(ss), npix, C, st,))
    return a, mean, var


def bn_train_bwd_nhwc_bf16(da, y, mean, var, gamma, beta, eps, relu, dg_out=None, db_out=None):
    """-> (dy bf16, dgamma, dbeta); dg_out / db_out: fp32 [C] tensors to write the parameter gradients into."""
    _bf16(da, "da"); _bf16(y, "y"); _f32(dg_out, "dg_out"); _f32(db_out, "db_out")
    C = y.shape[-1]
    npix = y.numel() // C
    dy = torch.empty_like(y)
    dg = torch.empty(C, dtype=torch.float32, device=y.device) if dg_out is None else dg_out
    db = torch.empty(C, dtype=torch.float32, device=y.device) if db_out is None else db_out
    ws = _red_ws(2 * C, y.device)
    coef = torch.empty(4 * C, dtype=torch.float32, device=y.device)
    lib, st = _prep(da, y, mean, var, gamma, beta, dy, dg, db, ws, coef)
    _launch(lib, "pmu_bn_train_bwd_nhwc_bf16", (_p(da), _p(y), _p(mean), _p(var), _p(gamma), _p(beta), float(eps), int(relu),
                                              _p(dy), _p(dg), _p(db), _p(ws), _p(coef), npix, C, st,))
    return dy, dg, db


def channel_sums_nhwc_bf16(x, per_image=False):
    """Sum over the pixels of a bf16 NHWC tensor: [C], or [B, C] (one row per image) with per_image."""
    _bf16(x, "x")
    C = x.shape[-1]
    nseg = x.shape[0] if per_image else 1
    out = torch.empty((nseg, C) if per_image else (C,), dtype=torch.float32, device=x.device)
    ws = _red_ws(C, x.device)
    lib, st = _prep(x, out, ws)
    _launch(lib, "pmu_channel_sums_nhwc_bf16", (_p(x), _p(out), _p(ws), nseg, x.numel() // (C * nseg), C, st,))
    return out


def pool2_bwd_nhwc_bf16(x, dy, mode, in_hw=None):
    """x: the pooling input [B,H,W,C] (None for the average, then in_hw = (H, W)); dy [B,Ho,Wo,C] -> dx [B,H,W,C]."""
    _bf16(x, "x"); _bf16(dy, "dy")
    B, Ho, Wo, C = dy.shape
    H, W = (x.shape[1], x.shape[2]) if x is not None else in_hw
    dx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dy.device)
    lib, st = _prep(x, dy, dx)
    _launch(lib, "pmu_pool2_bwd_nhwc_bf16", (_p(x), _p(dy), _p(dx), B, H, W, C, mode, st,))
    return dx


def add_bf16_(dst, src):
    _bf16(dst, "dst"); _bf16(src, "src")
    assert dst.shape == src.shape
    lib, st = _prep(dst, src)
    _launch(lib, "pmu_add_bf16", (_p(dst), _p(src), dst.numel(), st,))
    return dst


def gauss_head_bwd_nhwc_bf16(enc, w, dmu, dls, dw, db):
    _bf16(enc, "enc")
    B, h, w_, C = enc.shape
    denc = torch.empty_like(enc)
    lib, st = _prep(enc, w, dmu, dls, denc, dw, db)
    _launch(lib, "pmu_gauss_head_bwd_nhwc_bf16", (_p(enc), _p(w), _p(dmu), _p(dls), _p(denc), _p(dw), _p(db), B, C, h, w_,
                                                dmu.shape[1], st,))
    return denc


def pack_conv3x3_weights_bf16(w, want_fwd=True, want_dgrad=True):
    """fp32 OIHW conv weights -> (wf bf16 [Cout, 9*Cin] forward operand, wd bf16 [Cin, 9*Cout] data-gradient operand with
    the taps flipped), one kernel."""
    _f32(w, "w")
    Cout, Cin = w.shape[0], w.shape[1]
    wf = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device=w.device) if want_fwd else None
    wd = torch.empty(Cin, 9 * Cout, dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    lib, st = _prep(w, wf, wd)
    _launch(lib, "pmu_pack_conv3x3_weights_bf16", (_p(w), _p(wf), _p(wd), Cout, Cin, st,))
    return wf, wd


def unpack_conv3x3_wgrad_f32(dwp, out=None):
    """fp32 [Cout, 9, Cin] (pmu_conv_wgrad_bf16) -> OIHW [Cout, Cin, 3, 3]."""
    _f32(dwp, "dwp")
    Cout, _, Cin = dwp.shape
    dw = torch.empty(Cout, Cin, 3, 3, dtype=torch.float32, device=dwp.device) if out is None else out
    lib, st = _prep(dwp, dw)
    _launch(lib, "pmu_unpack_conv3x3_wgrad_f32", (_p(dwp), _p(dw), Cout, Cin, st,))
    return dw


def conv1x1_slicebias_bf16(x, wpack, bias, relu):
    """tcgen05 1x1 convolution with a per-image bias [B, Cout]: x 16-bit NHWC -> y NHWC in the same format."""
    _f32(bias, "bias")
    wf = _h16((x, "x"), (wpack, "wpack"))
    B, H, W, Cin = x.shape
    Cout = wpack.shape[0]
    assert tuple(bias.shape) == (B, Cout), bias.shape
    y = torch.empty(B, H, W, Cout, dtype=x.dtype, device=x.device)
    lib, st = _prep(x, wpack, bias, y)
    global _META
    _META = {"flops": 2.0 * B * H * W * Cout * Cin}
    _launch(lib, "pmu_conv1x1_slicebias_bf16", (_p(x), Cin, _p(wpack), _p(bias), _p(y), B, H, W, Cout, int(relu), wf, st,))
    return y


def fcomb_last_fwd_bf16(h, w, bias):
    """h bf16 [B,H,W,F], w fp32 [C,F], bias [C] -> logits fp32 [B,C,H,W]."""
    _bf16(h, "h"); _f32(w, "w")
    B, H, W, F = h.shape
    C = w.shape[0]
    logits = torch.empty(B, C, H, W, dtype=torch.float32, device=h.device)
    lib, st = _prep(h, w, bias, logits)
    _launch(lib, "pmu_fcomb_last_fwd_bf16", (_p(h), _p(w), _p(bias), _p(logits), B, H * W, F, C, st,))
    return logits


def fcomb_last_bwd_bf16(h, dlogits, w):
    """-> (dh bf16 [B,H,W,F] with the ReLU mask of h applied, dw fp32 [C,F])."""
    _bf16(h, "h"); _f32(dlogits, "dlogits"); _f32(w, "w")
    B, H, W, F = h.shape
    C = w.shape[0]
    dh = torch.empty_like(h)
    dw = torch.empty(C, F, dtype=torch.float32, device=h.device)
    ws = _red_ws(C * F, h.device)
    lib, st = _prep(h, dlogits, w, dh, dw, ws)
    _launch(lib, "pmu_fcomb_last_bwd_bf16", (_p(h), _p(dlogits), _p(w), _p(dh), _p(dw), _p(ws), B, H * W, F, C, st,))
    return dh, dw


def relu_mask_bf16_(d, h):
    _bf16(d, "d"); _bf16(h, "h")
    assert d.shape == h.shape
    lib, st = _prep(d, h)
    _launch(lib, "pmu_relu_mask_bf16", (_p(d), _p(h), d.numel(), st,))
    return d


def conv3x3_wgrad_smallcin_bf16(x0, dy, x1=None, out=None):
    """First-layer weight gradient: x0 (and x1) fp32 [B,1,H,W], dy bf16 [B,H,W,Cout] -> dw fp32 [Cout, 1 or 2, 3, 3]."""
    _f32(x0, "x0"); _f32(x1, "x1"); _bf16(dy, "dy"); _f32(out, "out")
    B, H, W, Cout = dy.shape
    assert x0.shape[1] == 1 and (x1 is None or x1.shape[1] == 1)
    dw = torch.empty(Cout, 1 if x1 is None else 2, 3, 3, dtype=torch.float32, device=dy.device) if out is None else out
    ws = _red_ws(9 * Cout, dy.device)
    lib, st = _prep(x0, x1, dy, dw, ws)
    _launch(lib, "pmu_conv3x3_wgrad_smallcin_bf16", (_p(x0), _p(x1), _p(dy), _p(dw), _p(ws), B, H, W, Cout, st,))
    return dw


def conv_gemm_bnstats_bf16(x0, wpack, bias, Cout, ntaps, stats, x1=None):
    """tcgen05 conv (no ReLU) whose epilogue also accumulates the BatchNorm batch statistics of its output:
    stats fp64 [2*Cout] (zero-filled by the caller) += per-channel {sum, sum of squares} interleaved.  Returns y."""
    _f32(bias, "bias")
    wf = _h16((x0, "x0"), (x1, "x1"), (wpack, "wpack"))
    if stats.dtype != torch.float64 or stats.numel() != 2 * Cout:
        raise RuntimeError("conv_gemm_bnstats_bf16: stats must be float64 [2 * Cout]")
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    out = torch.empty(B, H, W, Cout, dtype=x0.dtype, device=x0.device)
    lib, st = _prep(x0, x1, wpack, bias, out, stats)
    global _META
    _META = {"flops": 2.0 * B * H * W * Cout * (9 if ntaps == 9 else 1) * (C0 + C1)}
    _launch(lib, "pmu_conv_gemm_bnstats_bf16", (_p(x0), C0, _p(x1), C1, _p(wpack), _p(bias), _p(out), _p(stats), B, H, W, Cout,
                                              int(ntaps), wf, st,))
    return out


def bn_train_fwd_stats_nhwc_bf16(y, stats, gamma, beta, eps, relu, momentum=0.1, run_mean=None, run_var=None):
    """Train-mode BatchNorm2d (+ReLU) on y bf16 [B,H,W,C] from the statistics the convolution accumulated: (a, mean, var)."""
    _bf16(y, "y")
    C = y.shape[-1]
    npix = y.numel() // C
    mean = torch.empty(C, dtype=torch.float32, device=y.device)
    var = torch.empty_like(mean)
    a = torch.empty_like(y)
    ss = torch.empty(2 * C, dtype=torch.float32, device=y.device)
    lib, st = _prep(y, stats, gamma, beta, run_mean, run_var, mean, var, a, ss)
    _launch(lib, "pmu_bn_train_fwd_stats_nhwc_bf16", (_p(y), _p(stats), _p(gamma), _p(beta), float(eps), int(relu), float(momentum),
                                                    _p(run_mean), _p(run_var), _p(mean), _p(var), _p(a), _p(ss), npix, C, st,))
    return a, mean, var


class PackedConvWeights:
    """Persistent bf16 operand copies (forward + data-gradient layouts) of a list of fp32 OIHW 3x3 conv weights, refreshed
    by ONE kernel launch (`refresh()`); `get(w)` -> (wf [Cout, 9*Cin], wd [Cin, 9*Cout])."""

    def __init__(self, weights):
        self.weights = [w for w in weights]
        dev = self.weights[0].device
        total = sum(2 * w.numel() for w in self.weights)
        self.buf = torch.empty(total, dtype=torch.bfloat16, device=dev)
        rows, self.views, off, tile = [], {}, 0, 0
        for w in self.weights:
            _f32(w, "w")
            Cout, Cin = int(w.shape[0]), int(w.shape[1])
            if w.shape[2:] != (3, 3) or Cout % 32 or Cin % 32 or not w.is_contiguous():
                raise RuntimeError("PackedConvWeights: contiguous [Cout, Cin, 3, 3] weights with channels % 32 == 0")
            n = w.numel()
            wf = self.buf[off:off + n].view(Cout, 9 * Cin)
            wd = self.buf[off + n:off + 2 * n].view(Cin, 9 * Cout)
            off += 2 * n
            rows.append([w.data_ptr(), wf.data_ptr(), wd.data_ptr(), Cout, Cin, tile])
            tile += (Cout // 32) * (Cin // 32)
            self.views[w.data_ptr()] = (wf, wd)
        self.total_tiles = tile
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.key = tuple(w.data_ptr() for w in self.weights)

    def matches(self, weights) -> bool:
        return self.key == tuple(w.data_ptr() for w in weights)

    def refresh(self):
        lib, st = _prep(self.buf, self.table)
        _launch(lib, "pmu_pack_conv3x3_weights_multi_bf16", (_p(self.table), len(self.weights), self.total_tiles, st,))

    def get(self, w):
        return self.views.get(w.data_ptr())
