"""Training step of ProbabilisticUnet through the CUDA kernels: fp32 NCHW parity mode and the bf16
tensor-core mode (precision="bf16" / "f16": every GEMM of the step — convolution forward / dgrad / wgrad, the transposed
convolutions, the Fcomb 1x1 chain — on tcgen05, activations and their gradients bf16 NHWC END TO END: BatchNorm (its batch
statistics come out of the convolution epilogue) / pooling / skip adds / the Gaussian head run on that layout
(csrc/train_bf16.cu), so no layout or precision cast sits between two GEMMs of the step; gradients land in one flat buffer
in completion order, which the data-parallel CUDA-graph step all-reduces range by range under the backward).

What the reference's training loop asks of autograd (train.py:85-110 via
ProbUNetTrainer.predict / loss, trainer/probunet_trainer.py:27-39):

    net.forward(imgs, masks, training=True)      # U-Net, prior, posterior — train-mode BatchNorm
    loss = -net.elbo(masks)                       # z_q = rsample, fcomb, sum CE + beta * mean KL
    loss.backward()                               # gradients of every parameter

Here forward records what the backward kernels need (conv inputs, pre-BN outputs, batch
statistics), and ``elbo`` returns a scalar attached to ONE autograd node (`_ElboFn`) whose
backward runs the explicit backward kernels of csrc/train_f32.cu and hands each parameter its
gradient — so ``loss.backward()``, gradient accumulation over ``acc_steps``,
``clip_grad_value_`` and ``optim.SGD`` of train.py work unchanged.  No ATen kernel computes
any of the arithmetic; torch owns memory and the [B, L]-sized reparameterisation glue.

Layer backward formulas follow the torch ops the reference uses (nn.Conv2d, BatchNorm2d in
training mode, ReLU, MaxPool2d(2), AvgPool2d(2,2,ceil_mode), ConvTranspose2d(k2,s2), torch.mean,
CrossEntropyLoss(reduction none -> sum), kl_divergence of Independent Normals).
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


def _d(t):
    return t.detach()


# per-net state of the tensor-core step (completion order of the gradients, flat layout, persistent packed weights): kept
# here, weakly keyed by the module, so that nothing of it is pickled, deep-copied or saved with the model
_NET_STATE = weakref.WeakKeyDictionary()


def _net_state(net) -> dict:
    st = _NET_STATE.get(net)
    if st is None:
        st = _NET_STATE[net] = {}
    return st


class _FlatLayout:
    """Where each parameter's gradient lives in the ONE fp32 buffer of a backward: parameters in the order the backward
    completes them (recorded from a previous backward of the same net), so that a finished prefix is a contiguous range —
    the data-parallel step all-reduces range after range while the rest of the backward still runs, with no flatten /
    unflatten copies (train_dp.py)."""

    def __init__(self, order: List[nn.Parameter]):
        self.slots: Dict[int, tuple] = {}
        n = 0
        for p in order:
            self.slots[id(p)] = (n, p.numel(), tuple(p.shape))
            n += (p.numel() + 3) // 4 * 4               # views stay 16-byte aligned
        self.total = n
        self.order = [id(p) for p in order]


class _Tape:
    """Gradients by parameter (id -> tensor); each parameter is written at most once per step.

    With a layout: the gradients are views of `buf` (zero-filled); kernels write them in place through `out(p)`, anything
    else is copied in by `put`.  `on_ready(buf, lo, hi)` is called whenever the completed prefix has grown by
    `bucket_bytes` (and once at `finish()`): the parameters of [lo, hi) are final."""

    def __init__(self, layout: Optional[_FlatLayout] = None, device=None, on_ready=None, bucket_bytes: int = 32 << 20):
        self.g: Dict[int, torch.Tensor] = {}
        self.order: List[nn.Parameter] = []
        self.layout = layout
        self.buf = torch.zeros(layout.total, dtype=torch.float32, device=device) if layout is not None else None
        self.on_ready, self.bucket = on_ready, bucket_bytes // 4
        self.done_lo = self.done_hi = 0
        self.in_order = True
        self.npos = 0

    def out(self, p: nn.Parameter) -> Optional[torch.Tensor]:
        """The (zero-filled) place of p's gradient in the flat buffer, for a kernel to write; None without a layout."""
        if self.layout is None:
            return None
        off, n, shape = self.layout.slots[id(p)]
        return self.buf[off:off + n].view(shape)

    def put(self, p: nn.Parameter, g: Optional[torch.Tensor] = None):
        """g = the gradient; None = it already sits in out(p) (written in place, or exactly zero)."""
        if id(p) in self.g:
            ops.add_f32_(self.g[id(p)], g.reshape(p.shape).contiguous())
            return
        self.order.append(p)
        if self.layout is None:
            self.g[id(p)] = g.reshape(p.shape)
            return
        v = self.out(p)
        if g is not None and g.data_ptr() != v.data_ptr():
            v.copy_(g.reshape(p.shape))
        self.g[id(p)] = v
        # completed prefix (only meaningful while the backward follows the recorded order)
        if self.in_order and self.npos < len(self.layout.order) and self.layout.order[self.npos] == id(p):
            self.npos += 1
            off, n, _ = self.layout.slots[id(p)]
            self.done_hi = off + (n + 3) // 4 * 4
            if self.on_ready is not None and self.done_hi - self.done_lo >= self.bucket:
                self.on_ready(self.buf, self.done_lo, self.done_hi)
                self.done_lo = self.done_hi
        else:
            self.in_order = False

    def finish(self):
        if self.layout is not None and self.on_ready is not None and self.done_lo < self.layout.total:
            self.on_ready(self.buf, self.done_lo, self.layout.total)
            self.done_lo = self.layout.total


# ------------------------------------------------------------------ conv3x3 + BN(train) + ReLU
# Self-check hook (tests / debugging): when CHECK_LOG is a list, every tcgen05 forward / dgrad / wgrad of a tensor-core step
# is recomputed by the fp32 CUDA-core kernel on the SAME (bf16-rounded) operands and (kind, Cin, Cout, H, relative L2
# deviation) is appended — the deviation is then accumulation order + the bf16 rounding of the GEMM's output (~3e-3).
CHECK_LOG = None
_BF16 = {"on": False, "arena": None}
_PENDING_NBT: List[torch.Tensor] = []     # BatchNorm num_batches_tracked counters of this forward: ONE foreach add at its end


class _Arena:
    """Zero-filled fp32 scratch of one backward: the small accumulators kernels add into and the zero bias gradients are
    views of ONE buffer cleared by ONE fill (larger requests fall back to torch.zeros)."""

    def __init__(self, n: int, device):
        self.buf = torch.zeros(n, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        if self.off + n > self.buf.numel():
            return torch.zeros(*shape, dtype=torch.float32, device=self.buf.device)
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 3) // 4 * 4                  # views stay 16-byte aligned
        return v


FUSED_BN_STATS = True      # tensor-core mode: BatchNorm statistics in the convolution epilogue (False: the separate statistics pass)


def _stats_take(C: int, device) -> Optional[torch.Tensor]:
    """fp64 [2*C] zero-filled accumulator for the statistics of one conv + BatchNorm layer, cut from ONE buffer per forward."""
    if not FUSED_BN_STATS:
        return None
    st = _BF16.get("stats")
    if st is None or st[1] + 2 * C > st[0].numel():
        st = _BF16["stats"] = [torch.zeros(max(1 << 16, 2 * C), dtype=torch.float64, device=device), 0]
    v = st[0][st[1]:st[1] + 2 * C]
    st[1] += 2 * C
    return v


def _zeros(*shape, device=None):
    a = _BF16["arena"]
    return a.take(*shape) if a is not None else torch.zeros(*shape, dtype=torch.float32, device=device)


def _cbr_fwd(conv: nn.Conv2d, bn: nn.BatchNorm2d, x0, x1=None):
    if _BF16["on"]:
        return _tc_cbr_fwd(conv, bn, x0, x1)
    y = ops.conv3x3_f32(x0, _d(conv.weight), _d(conv.bias), relu=False, x1=x1)
    a, mean, var = ops.bn_train_fwd_f32(y, _d(bn.weight), _d(bn.bias), bn.eps, True,
                                        0.1 if bn.momentum is None else bn.momentum, bn.running_mean, bn.running_var)
    _PENDING_NBT.append(bn.num_batches_tracked)
    return a, {"conv": conv, "bn": bn, "x0": x0, "x1": x1, "y": y, "mean": mean, "var": var}


def _cbr_bwd(rec, da, tape: _Tape, need_dx=True):
    if _BF16["on"]:
        return _tc_cbr_bwd(rec, da, tape, need_dx)
    conv, bn = rec["conv"], rec["bn"]
    dy, dg, db, dcb = ops.bn_train_bwd_f32(da, rec["y"], rec["mean"], rec["var"], _d(bn.weight), _d(bn.bias), bn.eps, True,
                                           want_dbias=True)
    tape.put(bn.weight, dg)
    tape.put(bn.bias, db)
    tape.put(conv.bias, dcb)
    w = _d(conv.weight)
    C0 = rec["x0"].shape[1]
    dw = torch.zeros_like(w)
    ops.conv3x3_wgrad_f32(rec["x0"], dy, dw, rec["x1"])
    tape.put(conv.weight, dw)
    if not need_dx:
        return None, None
    # data gradient = the same 3x3 kernel with the weights transposed (ci <-> co) and flipped
    wt = w.flip(2, 3).transpose(0, 1)
    dx0 = ops.conv3x3_f32(dy, wt[:C0].contiguous(), None, relu=False)
    dx1 = ops.conv3x3_f32(dy, wt[C0:].contiguous(), None, relu=False) if rec["x1"] is not None else None
    return dx0, dx1


# ------------------------------------------------------------------ the same layer in the tensor-core mode
# Activations and gradients are bf16 NHWC [B,H,W,C].  The three first layers (Cin = 1: U-Net, prior; Cin = 2: posterior) take the
# fp32 NCHW image (+ mask) and have no data gradient; their weight gradient (576 / 1152 values) uses the fp32 small-Cin
# kernel on a cast of dy — the only casts left in the step besides the fcomb head's input / output.
def _f32(t_bf16_nhwc):
    return ops.nhwc_bf16_to_nchw_f32(t_bf16_nhwc)


def _tc_cbr_fwd(conv, bn, x0, x1=None):
    w = _d(conv.weight)
    Cout, Cin = w.shape[0], w.shape[1]
    first = Cin <= 2
    wd = None
    if first:
        if Cout % 8 or Cout > 128:
            raise NotImplementedError("tensor-core training needs a first layer with Cout % 8 == 0, <= 128")
        y = ops.conv3x3_first_bf16(x0, w, _d(conv.bias), relu=False, x1=x1, out_dtype=torch.bfloat16)
    else:
        # both operand layouts of this layer from one read of the fp32 weights: wf [Cout][tap][Cin] for this GEMM,
        # wd [Cin][flipped tap][Cout] for the data gradient in the backward
        pk = _BF16.get("packed")
        pre = pk.get(w) if pk is not None else None
        wf, wd = pre if pre is not None else ops.pack_conv3x3_weights_bf16(w)
        stats = _stats_take(Cout, x0.device)
        if stats is not None:
            # BatchNorm batch statistics in the convolution's epilogue: the statistics pass over y is gone
            y = ops.conv_gemm_bnstats_bf16(x0, wf, _d(conv.bias), Cout, 9, stats, x1=x1)
        else:
            y = ops.conv_gemm_bf16(x0, wf, _d(conv.bias), Cout, 9, False, x1=x1)
        if CHECK_LOG is not None:
            ref = ops.conv3x3_f32(_f32(x0), w.to(torch.bfloat16).float(), _d(conv.bias), relu=False, x1=None if x1 is None else _f32(x1))
            CHECK_LOG.append(("fwd", Cin, Cout, y.shape[1], float((_f32(y) - ref).norm() / ref.norm())))
    mom = 0.1 if bn.momentum is None else bn.momentum
    if not first and stats is not None:
        a, mean, var = ops.bn_train_fwd_stats_nhwc_bf16(y, stats, _d(bn.weight), _d(bn.bias), bn.eps, True, mom,
                                                        bn.running_mean, bn.running_var)
    else:
        a, mean, var = ops.bn_train_fwd_nhwc_bf16(y, _d(bn.weight), _d(bn.bias), bn.eps, True, mom, bn.running_mean, bn.running_var)
    _PENDING_NBT.append(bn.num_batches_tracked)
    return a, {"conv": conv, "bn": bn, "x0": x0, "x1": x1, "y": y, "mean": mean, "var": var, "first": first, "wd": wd}


def _tc_cbr_bwd(rec, da, tape: _Tape, need_dx=True):
    conv, bn = rec["conv"], rec["bn"]
    dy, dg, db = ops.bn_train_bwd_nhwc_bf16(da, rec["y"], rec["mean"], rec["var"], _d(bn.weight), _d(bn.bias), bn.eps, True,
                                            dg_out=tape.out(bn.weight), db_out=tape.out(bn.bias))
    tape.put(bn.weight, dg)
    tape.put(bn.bias, db)
    # a bias in front of a train-mode BatchNorm has an exactly zero gradient (BatchNorm subtracts the batch mean); autograd's
    # value is the rounding residue of sum(dy), which the fp32 path reproduces and a bf16 dy would only replace by other noise
    tape.put(conv.bias, None if tape.layout is not None else _zeros(conv.bias.numel(), device=dy.device))
    w = _d(conv.weight)
    Cout, Cin = w.shape[0], w.shape[1]
    if rec["first"]:
        tape.put(conv.weight, ops.conv3x3_wgrad_smallcin_bf16(rec["x0"], dy, rec["x1"], out=tape.out(conv.weight)))
        return None, None
    C0 = rec["x0"].shape[3]
    dwp = torch.empty(Cout, 9, Cin, dtype=torch.float32, device=dy.device)
    ops.conv_wgrad_bf16(rec["x0"], dy, dwp, rec["x1"], 9, overwrite=True)
    tape.put(conv.weight, ops.unpack_conv3x3_wgrad_f32(dwp, out=tape.out(conv.weight)))
    if CHECK_LOG is not None:
        ref = torch.zeros_like(w)
        ops.conv3x3_wgrad_f32(_f32(rec["x0"]), _f32(dy), ref, None if rec["x1"] is None else _f32(rec["x1"]))
        got = dwp.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
        CHECK_LOG.append(("wgrad", Cin, Cout, dy.shape[1], float((got - ref).norm() / ref.norm())))
    if not need_dx:
        return None, None
    # data gradient: the forward tcgen05 kernel with W transposed (ci <-> co) and flipped, [ci][tap][co] (packed in the forward)
    wt = rec["wd"]
    dx0 = ops.conv_gemm_bf16(dy, wt[:C0], None, C0, 9, False)
    if CHECK_LOG is not None:
        ref = ops.conv3x3_f32(_f32(dy), w.to(torch.bfloat16).float().flip(2, 3).transpose(0, 1)[:C0].contiguous(), None, relu=False)
        CHECK_LOG.append(("dgrad", Cin, Cout, dy.shape[1], float((_f32(dx0) - ref).norm() / ref.norm())))
    dx1 = ops.conv_gemm_bf16(dy, wt[C0:], None, Cin - C0, 9, False) if rec["x1"] is not None else None
    return dx0, dx1


def _dconv_fwd(dc, x0, x1=None):
    seq = dc.double_conv
    a, r1 = _cbr_fwd(seq[0], seq[1], x0, x1)
    b, r2 = _cbr_fwd(seq[3], seq[4], a)
    return b, (r1, r2)


def _dconv_bwd(recs, d, tape, need_dx=True):
    d, _ = _cbr_bwd(recs[1], d, tape)
    return _cbr_bwd(recs[0], d, tape, need_dx)


# ------------------------------------------------------------------ U-Net (unet_model.py:31-54)
def _tc() -> bool:
    return _BF16["on"]


def _pool_fwd(h, mode):
    return ops.pool2_bf16(h, mode) if _tc() else ops.pool2_f32(h, mode)


def _hw(t):
    return (t.shape[1], t.shape[2]) if _tc() else (t.shape[2], t.shape[3])


def _unet_fwd(unet, x):
    h, r = _dconv_fwd(unet.inc, x)
    skips, recs = [h], [r]
    for down in unet.down_blocks:
        p = _pool_fwd(h, ops.POOL_MAX)
        h, r = _dconv_fwd(down.maxpool_conv[1], p)
        skips.append(h)
        recs.append(r)
    ups = []
    n = len(unet.down_blocks)
    for i, up in enumerate(unet.up_blocks):
        skip = skips[n - 1 - i]
        if _hw(skip) != (2 * _hw(h)[0], 2 * _hw(h)[1]):
            raise NotImplementedError("training needs H, W divisible by 2^levels (the F.pad branch of Up.forward, "
                                      "unet_parts.py:58-62, is only built for inference)")
        wt = _d(up.up.weight)                                   # [Cin, Cout, 2, 2]
        if _tc():
            if wt.shape[0] % 64 or wt.shape[1] % 64:
                raise NotImplementedError("tensor-core training needs transposed-convolution channels that are multiples of 64")
            # transposed convolution on tcgen05: one K tap, N = 4 * Cout (the four output phases share one A load)
            wpack = wt.permute(2, 3, 1, 0).reshape(4 * wt.shape[1], wt.shape[0]).to(torch.bfloat16).contiguous()
            u = ops.conv_gemm_bf16(h, wpack, _d(up.up.bias), wt.shape[1], 4, False)
            if CHECK_LOG is not None:
                ref = ops.convt2x2_f32(_f32(h), wt.to(torch.bfloat16).float(), _d(up.up.bias))
                CHECK_LOG.append(("convt", wt.shape[0], wt.shape[1], u.shape[1], float((_f32(u) - ref).norm() / ref.norm())))
        else:
            u = ops.convt2x2_f32(h, wt, _d(up.up.bias))
        h_in = h
        h, r = _dconv_fwd(up.conv, skip, u)        # cat([skip, up]) as a two-source convolution
        ups.append({"up": up, "h_in": h_in, "recs": r})
    return h, {"skips": skips, "recs": recs, "ups": ups}


def _unet_bwd(unet, st, dfeat, tape):
    n = len(unet.down_blocks)
    dskip: List[Optional[torch.Tensor]] = [None] * (n + 1)
    d = dfeat
    add_ = ops.add_bf16_ if _tc() else ops.add_f32_
    for i in reversed(range(len(st["ups"]))):
        u = st["ups"][i]
        ds, du = _dconv_bwd(u["recs"], d, tape)
        dskip[n - 1 - i] = ds
        up = u["up"].up
        wt = _d(up.weight)                                      # [Cin, Cout, 2, 2]
        Cin, Co = wt.shape[0], wt.shape[1]
        if _tc():
            tape.put(up.bias, ops.channel_sums_nhwc_bf16(du))
            # transposed-convolution backward on tcgen05: space-to-depth of du turns both gradients into 1x1 GEMMs
            D = ops.s2d_nhwc_bf16(du)                           # [B, H, W, (i, j, co)]
            dwp = torch.empty(4 * Co, 1, Cin, dtype=torch.float32, device=du.device)
            ops.conv_wgrad_bf16(u["h_in"], D, dwp, None, 1, overwrite=True)     # dwp[(i,j,co)][ci] = sum_pix D * x
            dw = dwp.reshape(2, 2, Co, Cin).permute(3, 2, 0, 1).contiguous()
            tape.put(up.weight, dw)
            wd = wt.permute(0, 2, 3, 1).reshape(Cin, 4 * Co).to(torch.bfloat16).contiguous()       # [ci][(i,j,co)]
            d = ops.conv_gemm_bf16(D, wd, None, Cin, 1, False)
            if CHECK_LOG is not None:
                ref = torch.zeros_like(wt)
                ops.convt2x2_wgrad_f32(_f32(u["h_in"]), _f32(du), ref)
                CHECK_LOG.append(("convt_wgrad", Cin, Co, du.shape[1], float((dw - ref).norm() / ref.norm())))
                ref = ops.convt2x2_dgrad_f32(_f32(du), wt.to(torch.bfloat16).float())
                CHECK_LOG.append(("convt_dgrad", Cin, Co, du.shape[1], float((_f32(d) - ref).norm() / ref.norm())))
        else:
            tape.put(up.bias, ops.channel_sums_f32(du))
            dw = torch.zeros_like(wt)
            ops.convt2x2_wgrad_f32(u["h_in"], du, dw)
            tape.put(up.weight, dw)
            d = ops.convt2x2_dgrad_f32(du, wt)
    # d = gradient w.r.t. the deepest encoder map
    for lvl in range(n, 0, -1):
        if dskip[lvl] is not None:
            add_(d, dskip[lvl])
        d, _ = _dconv_bwd(st["recs"][lvl], d, tape)
        d = ops.pool2_bwd_nhwc_bf16(st["skips"][lvl - 1], d, ops.POOL_MAX) if _tc() else ops.pool2_bwd_f32(st["skips"][lvl - 1], d, ops.POOL_MAX)
    if dskip[0] is not None:
        add_(d, dskip[0])
    _dconv_bwd(st["recs"][0], d, tape, need_dx=False)


# ------------------------------------------------------------------ prior / posterior (probabilistic_unet.py:11-114)
def _gauss_fwd(net, x, segm=None):
    layers = net.encoder.layers
    nblk = len(net.num_filters)
    h, recs, pool_in = x, [], []
    for i in range(nblk):
        if i > 0:
            pool_in.append(h)
            h = _pool_fwd(h, ops.POOL_AVG_CEIL)
        a, r1 = _cbr_fwd(layers[7 * i], layers[7 * i + 1], h, segm if i == 0 else None)
        h, r2 = _cbr_fwd(layers[7 * i + 3], layers[7 * i + 4], a)
        recs.append((r1, r2))
    cl = net.conv_layer
    L = cl.weight.shape[0] // 2
    head = ops.gauss_head_bf16 if _tc() else ops.gauss_head_f32
    mu, ls = head(h, _d(cl.weight).reshape(2 * L, -1), _d(cl.bias), L)
    return mu, ls, {"enc": h, "recs": recs, "pool_in": pool_in}


def _gauss_bwd(net, st, dmu, dls, tape):
    cl = net.conv_layer
    L = cl.weight.shape[0] // 2
    C = st["enc"].shape[3] if _tc() else st["enc"].shape[1]
    dw, db = tape.out(cl.weight), tape.out(cl.bias)          # accumulated with atomics: the flat buffer is zero-filled
    if dw is None:
        dw, db = _zeros(2 * L, C, device=dmu.device), _zeros(2 * L, device=dmu.device)
    else:
        dw = dw.view(2 * L, C)
    head_bwd = ops.gauss_head_bwd_nhwc_bf16 if _tc() else ops.gauss_head_bwd_f32
    d = head_bwd(st["enc"], _d(cl.weight).reshape(2 * L, -1), dmu.contiguous(), dls.contiguous(), dw, db)
    tape.put(cl.weight, dw)
    tape.put(cl.bias, db)
    for i in reversed(range(len(st["recs"]))):
        d, _ = _dconv_bwd(st["recs"][i], d, tape, need_dx=i > 0)
        if i > 0:
            pin = st["pool_in"][i - 1]
            d = ops.pool2_bwd_nhwc_bf16(None, d, ops.POOL_AVG_CEIL, in_hw=_hw(pin)) if _tc() else ops.pool2_bwd_f32(pin, d, ops.POOL_AVG_CEIL)


# ------------------------------------------------------------------ fcomb (probabilistic_unet.py:155-181)
def _fcomb_convs(fc):
    return [m for m in fc.layers if isinstance(m, nn.Conv2d)]


def _fcomb_fwd(fc, feat, z):
    convs = _fcomb_convs(fc)
    F_ = convs[0].weight.shape[0]
    L = convs[0].weight.shape[1] - F_
    w0 = _d(convs[0].weight).reshape(F_, F_ + L)
    zb = ops.fcomb_zbias_f32(z, w0, _d(convs[0].bias))
    hs = [ops.conv1x1_bb_f32(feat, w0, F_ + L, zb, F_, F_, F_, True)]
    for c in convs[1:]:
        hs.append(ops.conv1x1_bb_f32(hs[-1], _d(c.weight).reshape(F_, F_), F_, _d(c.bias), 0, F_, F_, True))
    last = fc.last_layer
    C = last.weight.shape[0]
    logits = ops.conv1x1_bb_f32(hs[-1], _d(last.weight).reshape(C, F_), F_, _d(last.bias), 0, F_, C, False)
    return logits, {"feat": feat, "z": z, "hs": hs}


def _fcomb_bwd(fc, st, dlogits, tape):
    convs = _fcomb_convs(fc)
    F_ = convs[0].weight.shape[0]
    L = convs[0].weight.shape[1] - F_
    last = fc.last_layer
    C = last.weight.shape[0]
    hs = st["hs"]
    dw = torch.zeros(C, F_, dtype=torch.float32, device=dlogits.device)
    ops.conv1x1_wgrad_f32(hs[-1], dlogits, dw)
    tape.put(last.weight, dw)
    tape.put(last.bias, ops.channel_sums_f32(dlogits))
    d = ops.conv1x1_f32(dlogits, _d(last.weight).reshape(C, F_).t().contiguous(), None)
    d = ops.relu_bwd_f32(hs[-1], d)
    for j in range(len(convs) - 1, 0, -1):
        c = convs[j]
        dw = torch.zeros(F_, F_, dtype=torch.float32, device=d.device)
        ops.conv1x1_wgrad_f32(hs[j - 1], d, dw)
        tape.put(c.weight, dw)
        tape.put(c.bias, ops.channel_sums_f32(d))
        d = ops.conv1x1_f32(d, _d(c.weight).reshape(F_, F_).t().contiguous(), None)
        d = ops.relu_bwd_f32(hs[j - 1], d)
    # layer 0: weight [F, F+L] = [feature part | latent part]
    w0 = _d(convs[0].weight).reshape(F_, F_ + L)
    dw0 = torch.zeros(F_, F_ + L, dtype=torch.float32, device=d.device)
    db0 = torch.zeros(F_, dtype=torch.float32, device=d.device)
    ops.conv1x1_wgrad_f32(st["feat"], d, dw0, ldw=F_ + L)
    B = d.shape[0]
    rs = ops.row_sums_f32(d, B * F_)
    dz = ops.fcomb_zbias_bwd_f32(rs, st["z"], w0, dw0, db0)
    tape.put(convs[0].weight, dw0)
    tape.put(convs[0].bias, db0)
    dfeat = ops.conv1x1_f32(d, w0[:, :F_].t().contiguous(), None)
    return dfeat, dz


# ------------------------------------------------------------------ fcomb, tensor-core mode (bf16 NHWC hidden maps)
def _fcomb_tc_ok(fc, feat) -> bool:
    """The tcgen05 form of the Fcomb training chain: feature / hidden width a multiple of 64, <= 4 classes, at least
    one hidden 1x1 layer after the first, images of >= 128 pixels (a GEMM tile lies in one image: per-slice bias)."""
    convs = _fcomb_convs(fc)
    F_ = convs[0].weight.shape[0]
    return (feat.dtype == torch.bfloat16 and F_ % 64 == 0 and feat.shape[3] == F_ and fc.last_layer.weight.shape[0] <= 4
            and all(c.weight.shape[0] == F_ and c.weight.shape[1] == F_ for c in convs[1:])
            and feat.shape[1] * feat.shape[2] >= 128 and feat.shape[2] >= 16 and feat.shape[1] >= 8)


def _fcomb_fwd_tc(fc, feat, z):
    """feat bf16 NHWC [B,H,W,F]; z fp32 [B,L].  Layer 0 = features GEMM + per-slice latent bias (the tiled z of
    probabilistic_unet.py:167-176 never materialises); hidden maps stay bf16 NHWC; logits fp32 NCHW."""
    convs = _fcomb_convs(fc)
    F_ = convs[0].weight.shape[0]
    L = convs[0].weight.shape[1] - F_
    w0 = _d(convs[0].weight).reshape(F_, F_ + L)
    zb = ops.fcomb_zbias_f32(z, w0, _d(convs[0].bias))                                  # [B, F]
    w0f = w0[:, :F_].to(torch.bfloat16).contiguous()
    hs = [ops.conv1x1_slicebias_bf16(feat, w0f, zb, True)]
    for c in convs[1:]:
        hs.append(ops.conv_gemm_bf16(hs[-1], _d(c.weight).reshape(F_, F_).to(torch.bfloat16), _d(c.bias), F_, 1, True))
    last = fc.last_layer
    C = last.weight.shape[0]
    logits = ops.fcomb_last_fwd_bf16(hs[-1], _d(last.weight).reshape(C, F_).contiguous(), _d(last.bias))
    return logits, {"feat": feat, "z": z, "hs": hs, "w0f": w0f, "tc": True}


def _fcomb_bwd_tc(fc, st, dlogits, tape):
    convs = _fcomb_convs(fc)
    F_ = convs[0].weight.shape[0]
    L = convs[0].weight.shape[1] - F_
    last = fc.last_layer
    C = last.weight.shape[0]
    hs = st["hs"]
    dev = dlogits.device
    d, dwl = ops.fcomb_last_bwd_bf16(hs[-1], dlogits, _d(last.weight).reshape(C, F_).contiguous())   # ReLU of hs[-1] folded in
    tape.put(last.weight, dwl)
    tape.put(last.bias, ops.channel_sums_f32(dlogits))
    for j in range(len(convs) - 1, 0, -1):
        c = convs[j]
        dwp = tape.out(c.weight)
        dwp = torch.empty(F_, 1, F_, dtype=torch.float32, device=dev) if dwp is None else dwp.view(F_, 1, F_)
        ops.conv_wgrad_bf16(hs[j - 1], d, dwp, None, 1, overwrite=True)                 # dw[co][ci] = sum_pix d[co] * h[ci]
        tape.put(c.weight, dwp)
        tape.put(c.bias, ops.channel_sums_nhwc_bf16(d))
        d = ops.conv_gemm_bf16(d, _d(c.weight).reshape(F_, F_).t().to(torch.bfloat16).contiguous(), None, F_, 1, False)
        ops.relu_mask_bf16_(d, hs[j - 1])
    # layer 0: weight [F, F+L] = [feature part | latent part]
    w0 = _d(convs[0].weight).reshape(F_, F_ + L)
    dw0 = _zeros(F_, F_ + L, device=dev)
    db0 = _zeros(F_, device=dev)
    dwf = torch.empty(F_, 1, F_, dtype=torch.float32, device=dev)
    ops.conv_wgrad_bf16(st["feat"], d, dwf, None, 1, overwrite=True)
    rs = ops.channel_sums_nhwc_bf16(d, per_image=True)                   # [B, F]: what reaches the per-slice latent bias
    dz = ops.fcomb_zbias_bwd_f32(rs.reshape(-1), st["z"], w0, dw0, db0)
    dw0[:, :F_].copy_(dwf.reshape(F_, F_))
    tape.put(convs[0].weight, dw0)
    tape.put(convs[0].bias, db0)
    dfeat = ops.conv_gemm_bf16(d, st["w0f"].t().contiguous(), None, F_, 1, False)
    return dfeat, dz


def fcomb_forward(step, z):
    """Fcomb logits at latent z from a recorded training forward (net.sample() inside a training step)."""
    if step.tc_fcomb:
        return _fcomb_fwd_tc(step.net.fcomb, step.feat, z)[0]
    return _fcomb_fwd(step.net.fcomb, step.feat, z)[0]


def _prepack(net):
    """bf16 operand copies of every tensor-core 3x3 layer of the net for this step, one launch (ops.PackedConvWeights; the
    buffers persist on the net, so a CUDA graph sees static addresses)."""
    ws = [m.weight.detach() for part in (net.posterior, net.prior, net.unet) for m in part.modules()
          if isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.weight.shape[1] > 2
          and m.weight.shape[0] % 32 == 0 and m.weight.shape[1] % 32 == 0 and m.weight.is_contiguous()]
    if not ws:
        return None
    ns = _net_state(net)
    pk = ns.get("packed")
    if pk is None or not pk.matches(ws):
        pk = ns["packed"] = ops.PackedConvWeights(ws)
    pk.refresh()
    return pk


def _grad_layout(net) -> Optional[_FlatLayout]:
    """The flat-gradient layout from the completion order a previous backward recorded on `net` (None before the first
    one, or when a recorded parameter is no longer a trainable parameter of the net)."""
    ns = _net_state(net)
    order = ns.get("grad_order")
    if order is None:
        return None
    live = {id(p) for p in net.parameters() if p.requires_grad}
    if any(id(p) not in live for p in order):
        ns["grad_order"] = None
        return None
    lay = ns.get("grad_layout")
    if lay is None or lay.order != [id(p) for p in order]:
        lay = ns["grad_layout"] = _FlatLayout(order)
    return lay


# ------------------------------------------------------------------ the step
class TrainStep:
    """State of one forward(training=True) of a ProbabilisticUnet; consumed by elbo() / backward."""

    def __init__(self, net, patch: torch.Tensor, segm: torch.Tensor):
        self.net = net
        self.patch, self.segm = patch, segm
        self.bf16 = net.precision in ("bf16", "f16")      # the training GEMMs always take bf16 operands
        self._mode(True)
        try:
            self._forward(net, patch, segm)
        finally:
            self._mode(False)

    def _mode(self, on: bool):
        _BF16["on"] = on and self.bf16
        _BF16["stats"] = None
        if not on:
            _BF16["packed"] = None
            _BF16["arena"] = None
            _PENDING_NBT.clear()

    def _forward(self, net, patch, segm):
        if self.bf16:
            _BF16["packed"] = _prepack(net)
        self.mu_q, self.ls_q, self.post = _gauss_fwd(net.posterior, patch, segm)
        self.mu_p, self.ls_p, self.prior = _gauss_fwd(net.prior, patch)
        self.feat, self.unet = _unet_fwd(net.unet, patch)
        self.tc_fcomb = self.bf16 and _fcomb_tc_ok(net.fcomb, self.feat)
        if self.bf16 and not self.tc_fcomb:
            # shapes the tcgen05 Fcomb chain does not take: its fp32 NCHW kernels, one cast in, one out
            self.feat = ops.nhwc_bf16_to_nchw_f32(self.feat)
        self.fc = None
        if _PENDING_NBT:
            torch._foreach_add_(_PENDING_NBT, 1)
            _PENDING_NBT.clear()

    def features_nchw_f32(self) -> torch.Tensor:
        """The U-Net feature map as the reference exposes it (net.unet_features: fp32 NCHW)."""
        return ops.nhwc_bf16_to_nchw_f32(self.feat) if self.tc_fcomb else self.feat

    def _beta(self) -> float:
        # data parallel: the KL term averages over the GLOBAL batch (train_dp.py)
        return float(self.net.beta) / float(getattr(self.net, "kl_world_size", 1))

    def elbo(self, segm, z_q: torch.Tensor, eps: Optional[torch.Tensor], analytic_kl: bool):
        net = self.net
        if not analytic_kl:
            raise NotImplementedError("training backward is built for the analytic KL (the reference's default, "
                                      "probabilistic_unet.py:281)")
        self.z_q, self.eps_q = z_q.contiguous(), eps
        self.kl_b = ops.kl_diag_gauss(self.mu_q, self.ls_q, self.mu_p, self.ls_p)
        self.logits, self.fc = (_fcomb_fwd_tc if self.tc_fcomb else _fcomb_fwd)(net.fcomb, self.feat, self.z_q)
        self.segm_t = segm.contiguous().float()
        self.rec = ops.ce_sum(self.logits, self.segm_t)
        self.kl = self.kl_b.mean()
        return -(self.rec + self._beta() * self.kl)

    def backward(self, g: float, on_ready=None) -> Dict[int, torch.Tensor]:
        """g = d(loss)/d(elbo).  elbo = -(rec + beta * mean_b KL).  Tensor-core mode: the gradients are views of one flat
        buffer (a fresh one per backward), laid out in completion order once a previous backward of this net has recorded
        it; `on_ready(buf, lo, hi)` is then called for every completed range (the data-parallel all-reduce)."""
        self._on_ready = on_ready
        if self.unet is None:
            raise RuntimeError("this forward's activations were released by its first backward (one backward per "
                               "forward; retain_graph is not supported)")
        self._mode(True)
        try:
            return self._backward(g)
        finally:
            self._mode(False)
            # the recorded activations are dead now (one backward per forward, like autograd without retain_graph):
            # drop them at once instead of when the caller's loss tensor goes out of scope
            self.post = self.prior = self.unet = self.fc = self.logits = None

    def _backward(self, g: float) -> Dict[int, torch.Tensor]:
        net = self.net
        layout = _grad_layout(net) if self.bf16 else None
        tape = _Tape(layout, self.mu_q.device, getattr(self, "_on_ready", None))
        B = self.mu_q.shape[0]
        if self.bf16:
            # zero-filled scratch for the few accumulators that are added into (the weight gradients are write-only:
            # pmu_conv_wgrad_bf16(overwrite=1))
            _BF16["arena"] = _Arena(1 << 18, self.mu_q.device)
        dlogits = ops.ce_bwd_f32(self.logits, self.segm_t, -g)
        dfeat, dz = (_fcomb_bwd_tc if self.tc_fcomb else _fcomb_bwd)(net.fcomb, self.fc, dlogits, tape)
        dmu_q, dls_q, dmu_p, dls_p = ops.kl_bwd_f32(self.mu_q, self.ls_q, self.mu_p, self.ls_p, -g * self._beta() / B)
        if self.eps_q is not None:
            # z_q = mu_q + exp(log_sigma_q) * eps  (rsample): [B, L]-sized glue
            dmu_q = dmu_q + dz
            dls_q = dls_q + dz * self.eps_q * torch.exp(self.ls_q)
        _gauss_bwd(net.posterior, self.post, dmu_q, dls_q, tape)
        _gauss_bwd(net.prior, self.prior, dmu_p, dls_p, tape)
        _unet_bwd(net.unet, self.unet, ops.nchw_f32_to_nhwc_bf16(dfeat) if (self.bf16 and not self.tc_fcomb) else dfeat, tape)
        tape.finish()
        if self.bf16 and (layout is None or not tape.in_order or len(tape.order) != len(layout.order)):
            _net_state(net)["grad_order"] = list(tape.order)             # (re)record the completion order for the next backward
        self.flat = tape.buf
        return tape.g


class _ElboFn(torch.autograd.Function):
    """One autograd node for the whole step: inputs = every parameter, output = elbo."""

    @staticmethod
    def forward(ctx, step: TrainStep, value: torch.Tensor, *params):
        ctx.step = step
        ctx.params = params
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.step.backward(float(g))
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.params)


def elbo_with_grad(step: TrainStep, value: torch.Tensor, params):
    return _ElboFn.apply(step, value, *params)


# ------------------------------------------------------------------ the whole step as ONE CUDA graph
class GraphedTrainStep:
    """forward(training=True) + elbo + backward of a ProbabilisticUnet in train() mode, captured ONCE as a CUDA graph and
    replayed per step: the eager step issues ~3000 launches (2200 of this library + the weight re-packs) from Python and is
    bound by the host — 8.4 us per launch against 15.9 ms of kernel time per batch-8 step — the replay is one launch.

    Same kernels, same numbers as `net.forward(imgs, masks); loss = -net.elbo(masks); loss.backward()` (train.py:85-95): the
    graph holds static copies of the inputs, draws the posterior noise with torch's graph-safe generator and writes the
    gradients into static tensors, which `step()` hands to the parameters (`p.grad`), scaled for `acc_steps` like
    `loss / acc_steps`.  The out-of-range label check of the cross entropy (a host read-back) is made on the inputs at
    every `step()` unless `check_labels=False`."""

    def __init__(self, net, imgs: torch.Tensor, masks: torch.Tensor, loss_scale: float = 1.0, warmup: int = 3,
                 eps: Optional[torch.Tensor] = None, allreduce_group=None, allreduce: bool = False):
        if not (net.training and any(p.requires_grad for p in net.parameters())):
            raise RuntimeError("GraphedTrainStep needs net.train() and trainable parameters")
        if imgs.device.type != "cuda":
            raise RuntimeError("pmu_b200 models run on CUDA only (no CPU fallback)")
        self.net = net
        self.imgs = imgs.detach().clone().contiguous().float()
        self.masks = masks.detach().clone().contiguous().float()
        self.params = [p for p in net.parameters() if p.requires_grad]
        # injected posterior noise (tests: compare with the eager step on the same eps); None = drawn inside the graph
        self.eps = None if eps is None else eps.detach().clone().contiguous().float()
        self.loss_scale = float(loss_scale)
        self.kl_world = int(getattr(net, "kl_world_size", 1))
        # data parallel: SUM-all-reduce the flat gradient buffer INSIDE the graph, range by range as the backward completes
        # them — each collective is a side branch of the graph (NCCL's own stream, joined at the end), so the exchange of
        # the posterior's gradients runs under the prior's backward, the prior's under the U-Net's, and so on
        import torch.distributed as dist
        self.group = allreduce_group
        self.ar = bool(allreduce) and dist.is_available() and dist.is_initialized() and dist.get_world_size(allreduce_group) > 1
        self.ar_in_graph = False
        self.ar_ranges = []
        dev = imgs.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        ops.CE_CHECK_LABELS = False
        try:
            with torch.cuda.stream(side):                    # eager warm-up off the capture: module loads, allocator growth
                for _ in range(max(1, warmup)):
                    self._run()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self._run()
        finally:
            ops.CE_CHECK_LABELS = True

    def _run(self):
        net = self.net
        st = TrainStep(net, self.imgs, self.masks)
        eps = self.eps if self.eps is not None else torch.randn(st.mu_q.shape, dtype=torch.float32, device=self.imgs.device)
        z_q = st.mu_q + eps * torch.exp(st.ls_q)
        value = st.elbo(self.masks, z_q, eps, True)
        self.flag = ops.CE_LAST_FLAG
        works, ranges = [], []

        def exchange(buf, lo, hi):
            import torch.distributed as dist
            ranges.append((lo, hi))
            works.append(dist.all_reduce(buf[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

        grads = st.backward(-self.loss_scale, on_ready=exchange if self.ar else None)              # loss = -elbo * loss_scale
        for w in works:
            w.wait()                                       # joins NCCL's stream into the current one (a graph edge under capture)
        self.ar_in_graph, self.ar_ranges = bool(works) and st.flat is not None, ranges
        self.loss = -value * self.loss_scale
        self.kl, self.rec = st.kl, st.rec
        self.grads = [grads.get(id(p)) for p in self.params]

    def close(self):
        """Release the captured graph (and with it the NCCL operations recorded in it).  A data-parallel caller must do
        this before torch.distributed.destroy_process_group(): tearing down a communicator that a live CUDA graph still
        references does not return."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self.grads = []

    def step(self, imgs: torch.Tensor, masks: torch.Tensor, accumulate: bool = False, check_labels: bool = True,
             eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Replay on new inputs; gradients land in p.grad (added when `accumulate`).  Returns the loss (a static tensor:
        read it before the next step())."""
        if self.graph is None:
            raise RuntimeError("this GraphedTrainStep was closed")
        if tuple(imgs.shape) != tuple(self.imgs.shape) or tuple(masks.shape) != tuple(self.masks.shape):
            raise ValueError("GraphedTrainStep was captured for inputs of shape "
                             f"{tuple(self.imgs.shape)} / {tuple(self.masks.shape)}")
        if int(getattr(self.net, "kl_world_size", 1)) != self.kl_world:
            raise RuntimeError("the KL weighting (kl_world_size) changed since the capture")
        self.imgs.copy_(imgs, non_blocking=True)
        self.masks.copy_(masks, non_blocking=True)
        if eps is not None:
            if self.eps is None:
                raise ValueError("this graph draws its own posterior noise (captured without eps=)")
            self.eps.copy_(eps, non_blocking=True)
        self.graph.replay()
        if check_labels and self.flag is not None and float(self.flag) > 0:
            raise IndexError(f"Target out of bounds: labels must lie in [0, {self.net.n_classes}) (mask values outside the class range)")
        for p, g in zip(self.params, self.grads):
            if g is None:
                continue
            if accumulate and p.grad is not None:
                p.grad.add_(g)
            else:
                p.grad = g.clone() if accumulate else g
        return self.loss
