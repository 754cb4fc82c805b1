"""ctypes binding of libpmu_b200.so — the C-ABI declared in include/pmu_b200.h.

The product path has NO fallback: if the library is missing or a call fails, a
RuntimeError is raised with pmu_last_error().
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpmu_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "pmu_b200.h")

_lib = None

_P = c_void_p
_SIG = {
    "pmu_last_error": (c_char_p, []),
    "pmu_version": (c_int, []),
    "pmu_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pmu_set_device": (c_int, [c_int]),
    "pmu_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "pmu_ctx_destroy": (c_int, [c_void_p]),
    "pmu_ctx_bind": (c_int, [c_void_p]),
    "pmu_ctx_stats": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "pmu_plane_max": (c_int, [_P, POINTER(c_int32), _P, _P]),
    "pmu_slice_gather": (c_int, [_P, POINTER(c_int32), c_int, c_int, c_int, c_int, POINTER(c_float), c_int, c_int,
                                 _P, _P, _P, _P]),
    "pmu_slice_normalize": (c_int, [_P, _P, c_int, c_int64, _P]),
    "pmu_fill_f32": (c_int, [_P, c_float, c_int64, _P]),
    "pmu_conv3x3_f32": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_conv1x1_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int64, c_int, _P]),
    "pmu_convt2x2_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_pool2_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_gauss_head_f32": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_fcomb_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int64, _P]),
    "pmu_conv3x3_first_bf16": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_conv_gemm_bf16": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_conv_gemm_pool_bf16": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_pool2_bf16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_gauss_head_bf16": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_nhwc_bf16_to_nchw_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_fcomb_softmax_accum_bf16": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int,
                                             c_int, c_int64, c_int, _P]),
    "pmu_softmax_accum": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, _P]),
    "pmu_scatter_accum": (c_int, [_P, c_int, c_int, c_int, POINTER(c_int32), c_int, _P, _P, _P]),
    "pmu_bn_train_fwd_nhwc_bf16": (c_int, [_P, _P, _P, c_float, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "pmu_conv_gemm_bnstats_bf16": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_bn_train_fwd_stats_nhwc_bf16": (c_int, [_P, _P, _P, _P, c_float, c_int, c_float, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "pmu_bn_train_bwd_nhwc_bf16": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "pmu_channel_sums_nhwc_bf16": (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P]),
    "pmu_pool2_bwd_nhwc_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_add_bf16": (c_int, [_P, _P, c_int64, _P]),
    "pmu_gauss_head_bwd_nhwc_bf16": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_scatter_accum_affine": (c_int, [_P, POINTER(c_float), c_int, c_int, c_int, c_int, POINTER(c_int32), c_int, c_float, _P, _P, _P, _P]),
    "pmu_fuse_finalize_counted": (c_int, [_P, _P, _P, POINTER(c_int32), c_int, _P, _P, _P, _P, _P]),
    "pmu_fuse_finalize": (c_int, [_P, _P, c_float, POINTER(c_int32), c_int, _P, _P, _P, _P, _P]),
    "pmu_ce_sum": (c_int, [_P, _P, c_int, c_int, c_int64, _P, _P]),
    "pmu_kl_diag_gauss": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "pmu_dice_sums": (c_int, [_P, _P, c_int64, _P, _P]),
    "pmu_argmax_dice_sums": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P]),
    "pmu_bn_train_fwd_f32": (c_int, [_P, _P, _P, c_float, c_int, c_float, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int64, _P]),
    "pmu_bn_train_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P, c_int, c_int, c_int64, _P]),
    "pmu_bn_train_bwd_bias_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int64, _P]),
    "pmu_channel_sums_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int64, _P]),
    "pmu_row_sums_f32": (c_int, [_P, _P, c_int64, c_int64, _P]),
    "pmu_conv3x3_wgrad_f32": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "pmu_conv1x1_wgrad_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int64, _P]),
    "pmu_pool2_bwd_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_convt2x2_dgrad_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_convt2x2_wgrad_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_relu_bwd_f32": (c_int, [_P, _P, _P, c_int64, _P]),
    "pmu_add_f32": (c_int, [_P, _P, c_int64, _P]),
    "pmu_ce_bwd_f32": (c_int, [_P, _P, c_float, _P, c_int, c_int, c_int64, _P]),
    "pmu_kl_bwd_f32": (c_int, [_P, _P, _P, _P, c_float, _P, _P, _P, _P, c_int, c_int, _P]),
    "pmu_gauss_head_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_fcomb_zbias_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "pmu_fcomb_zbias_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "pmu_nchw_f32_to_nhwc_bf16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "pmu_s2d_nhwc_bf16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "pmu_conv_wgrad_bf16": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_pack_conv3x3_weights_bf16": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "pmu_pack_conv3x3_weights_multi_bf16": (c_int, [_P, c_int, c_int64, _P]),
    "pmu_unpack_conv3x3_wgrad_f32": (c_int, [_P, _P, c_int, c_int, _P]),
    "pmu_conv1x1_slicebias_bf16": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "pmu_fcomb_last_fwd_bf16": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P]),
    "pmu_fcomb_last_bwd_bf16": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P]),
    "pmu_conv3x3_wgrad_smallcin_bf16": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "pmu_relu_mask_bf16": (c_int, [_P, _P, c_int64, _P]),
    "pmu_conv1x1_bb_f32": (c_int, [_P, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int64, c_int, _P]),
}


def header_symbols(header_path: str = HEADER_PATH):
    """Every function name include/pmu_b200.h declares."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmu_[a-z0-9_]+)\s*\(", src)))


def load(build_if_missing: bool = False):
    """Load (once) and return the ctypes handle; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from .build import build_native
            build_native()
        else:
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                f"`python __graft_entry__.py build` (there is no CPU / eager fallback).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


USE_CTX = True        # False: the context-free launch path (pmu_set_device; every launch queries / encodes what it needs)
_ctx = {}


def bind_device(lib, device_index: int):
    """Make `device_index` current for the library on this thread: through its launch context (created on first use;
    cached TMA descriptors, kernel attributes, device properties), or through pmu_set_device when USE_CTX is off."""
    if not USE_CTX:
        check(lib.pmu_set_device(device_index), "pmu_set_device")
        return
    h = _ctx.get(device_index)
    if h is None:
        out = c_void_p()
        check(lib.pmu_ctx_create(device_index, ctypes.byref(out)), "pmu_ctx_create")
        h = _ctx[device_index] = out
    check(lib.pmu_ctx_bind(h), "pmu_ctx_bind")


def ctx_stats(device_index: int = 0):
    """(cached tensor maps, hits, misses) of the device's launch context; None before its first use."""
    h = _ctx.get(device_index)
    if h is None:
        return None
    a, b, c = c_int64(), c_int64(), c_int64()
    check(load().pmu_ctx_stats(h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "pmu_ctx_stats")
    return a.value, b.value, c.value


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pmu_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"pmu_b200 {what} failed (rc={rc}): {msg}")
