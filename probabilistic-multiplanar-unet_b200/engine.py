"""Network executor: packs a reference-schema state_dict once (BatchNorm folded, layouts
chosen per precision mode) and runs U-Net / prior / posterior / fcomb through the C-ABI ops.

Two precision modes (north star):
  "fp32": NCHW fp32, CUDA-core kernels          — parity mode (probabilities within 1e-4)
  "f16" : NHWC IEEE-half, tcgen05 implicit-GEMM convs — performance mode (within 2e-2; measured ~3e-3)
  "bf16": the same kernels on bfloat16 operands (the training step's format; per-view worst pixel 2.0-2.3e-2)

Reference call sites replaced: UNet.forward (model/unet/unet_model.py:31-54),
Encoder.forward / AxisAlignedConvGaussian.forward (probabilistic_unet.py:50-114),
Fcomb.forward (probabilistic_unet.py:167-181).  Eval-mode BatchNorm only (the canonical
inference path): BN is folded into the conv at pack time.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops

BN_EPS = 1e-5
# 16-bit storage format of each tensor-core mode (activations and packed weights alike; tcgen05 kind::f16 runs both at
# the same rate with fp32 accumulation).  "f16" (IEEE half, 11 significand bits) is the inference mode: bf16 operands
# put the worst pixel of one view's N-sample mean at 2.0-2.3e-2 on this 22-layer network — outside the 2e-2 bound —
# f16 operands at ~3e-3 (tests/tools/emulate_bf16_net.py).  "bf16" is what the training step uses and stays selectable.
H16 = {"f16": torch.float16, "bf16": torch.bfloat16}


def _fold(sd, conv: str, bn: str, dev):
    """conv+BN(eval) -> (w', b'):  w' = w*g/sqrt(v+eps),  b' = (b-mean)*g/sqrt(v+eps)+beta."""
    w = sd[conv + ".weight"].to(dev, torch.float32)
    b = sd[conv + ".bias"].to(dev, torch.float32)
    g = sd[bn + ".weight"].to(dev, torch.float32)
    beta = sd[bn + ".bias"].to(dev, torch.float32)
    mean = sd[bn + ".running_mean"].to(dev, torch.float32)
    var = sd[bn + ".running_var"].to(dev, torch.float32)
    s = g / torch.sqrt(var + BN_EPS)
    return (w * s[:, None, None, None]).contiguous(), ((b - mean) * s + beta).contiguous()


class _Conv:
    """One packed 3x3 conv (+folded BN): fp32 [Cout,Cin,3,3] and, in bf16 mode, [Cout][9][Cin] bf16."""

    def __init__(self, w, b, h16):
        self.w, self.b = w, b
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.tc = h16 is not None and self.cin % 64 == 0 and self.cout % 64 == 0
        self.first = h16 is not None and self.cin <= 2
        if h16 is not None and not (self.tc or self.first):
            raise RuntimeError(
                f"the tensor-core modes need channel counts that are multiples of 64 (conv {self.cin}->{self.cout}); "
                f"use precision='fp32' for this model")
        self.wpack = w.permute(0, 2, 3, 1).reshape(self.cout, 9 * self.cin).to(h16).contiguous() if self.tc else None


class PackedNet:
    """All weights of a ProbabilisticUnet, packed for one device + precision."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, precision: str = "fp32"):
        if precision not in ("fp32", "f16", "bf16"):
            raise ValueError(f"precision must be 'fp32', 'f16' or 'bf16', got {precision!r}")
        self.precision = precision
        self.device = torch.device(device)
        self.h16 = H16.get(precision)          # None in fp32 mode
        bf = self.h16
        dev = self.device
        L = 0
        while f"unet.down_blocks.{L}.maxpool_conv.1.double_conv.0.weight" in sd:
            L += 1
        self.levels = L + 1

        def dconv(p):
            return (_Conv(*_fold(sd, p + ".double_conv.0", p + ".double_conv.1", dev), bf),
                    _Conv(*_fold(sd, p + ".double_conv.3", p + ".double_conv.4", dev), bf))

        self.has_unet = "unet.inc.double_conv.0.weight" in sd
        if self.has_unet:
            self.inc = dconv("unet.inc")
            self.down = [dconv(f"unet.down_blocks.{i}.maxpool_conv.1") for i in range(L)]
            self.up = []
            for i in range(L):
                w = sd[f"unet.up_blocks.{i}.up.weight"].to(dev, torch.float32).contiguous()  # [Cin,Cout,2,2]
                b = sd[f"unet.up_blocks.{i}.up.bias"].to(dev, torch.float32).contiguous()
                cin, cout = w.shape[0], w.shape[1]
                wpack = None
                if bf is not None:
                    if cin % 64 or cout % 64:
                        raise RuntimeError(f"the tensor-core modes need convT channels multiple of 64 ({cin}->{cout})")
                    # rows = (i*2+j)*Cout + co, K = ci
                    wpack = w.permute(2, 3, 1, 0).reshape(4 * cout, cin).to(bf).contiguous()
                self.up.append({"w": w, "b": b, "wpack": wpack, "cout": cout, "conv": dconv(f"unet.up_blocks.{i}.conv")})
            if "unet.outc.conv.weight" in sd:
                self.outc_w = sd["unet.outc.conv.weight"].to(dev, torch.float32).reshape(
                    sd["unet.outc.conv.weight"].shape[0], -1).contiguous()
                self.outc_b = sd["unet.outc.conv.bias"].to(dev, torch.float32).contiguous()
        self.enc = {}
        for which in ("prior", "posterior"):
            if f"{which}.encoder.layers.0.weight" not in sd:
                continue
            layers = []
            i = 0
            while f"{which}.encoder.layers.{7 * i}.weight" in sd:
                p = f"{which}.encoder.layers."
                layers.append((_Conv(*_fold(sd, p + str(7 * i), p + str(7 * i + 1), dev), bf),
                               _Conv(*_fold(sd, p + str(7 * i + 3), p + str(7 * i + 4), dev), bf)))
                i += 1
            hw = sd[f"{which}.conv_layer.weight"].to(dev, torch.float32)
            self.enc[which] = {"layers": layers, "head_w": hw.reshape(hw.shape[0], -1).contiguous(),
                               "head_b": sd[f"{which}.conv_layer.bias"].to(dev, torch.float32).contiguous(),
                               "L": hw.shape[0] // 2}
        self.fcomb = None
        if "fcomb.layers.0.weight" in sd:
            w0 = sd["fcomb.layers.0.weight"].to(dev, torch.float32)
            F_ = w0.shape[0]
            mids_w, mids_b = [], []
            j = 1
            while f"fcomb.layers.{2 * j}.weight" in sd:
                mids_w.append(sd[f"fcomb.layers.{2 * j}.weight"].to(dev, torch.float32).reshape(F_, F_))
                mids_b.append(sd[f"fcomb.layers.{2 * j}.bias"].to(dev, torch.float32))
                j += 1
            wl = sd["fcomb.last_layer.weight"].to(dev, torch.float32)
            self.fcomb = {
                "w0": w0.reshape(F_, -1).contiguous(), "b0": sd["fcomb.layers.0.bias"].to(dev, torch.float32).contiguous(),
                "wmid": torch.stack(mids_w).contiguous() if mids_w else None,
                "bmid": torch.stack(mids_b).contiguous() if mids_b else None,
                "wlast": wl.reshape(wl.shape[0], -1).contiguous(),
                "blast": sd["fcomb.last_layer.bias"].to(dev, torch.float32).contiguous(),
                "nl": 2 + len(mids_w), "F": F_, "L": w0.shape[1] - F_, "C": wl.shape[0]}

    # ------------------------------------------------------------------ building blocks
    def _conv(self, c: _Conv, x, x1=None, first_x1=None):
        if self.precision == "fp32":
            return ops.conv3x3_f32(x, c.w, c.b, relu=True, x1=x1)
        if c.first:
            return ops.conv3x3_first_bf16(x, c.w, c.b, relu=True, x1=first_x1, out_dtype=self.h16)
        return ops.conv_gemm_bf16(x, c.wpack, c.b, c.cout, 9, True, x1=x1)

    def _pool(self, x, mode):
        return ops.pool2_f32(x, mode) if self.precision == "fp32" else ops.pool2_bf16(x, mode)

    # ------------------------------------------------------------------ networks
    def unet_features(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B,1,H,W] -> last decoder map: fp32 NCHW [B,F,H,W] or bf16 NHWC [B,H,W,F]."""
        fp32 = self.precision == "fp32"
        a, b = self.inc
        blocks = [self.inc] + list(self.down)
        xs, h = [], x
        for i, (a, b) in enumerate(blocks):
            last = i == len(blocks) - 1
            h = self._conv(a, h)
            if not last and not fp32 and b.tc and ops.fused_pool_ok(h.shape[1], h.shape[2]):
                # second conv of the block also emits the MaxPool2d(2) input of the next block
                full, h = ops.conv_gemm_pool_bf16(h, b.wpack, b.b, b.cout, True, ops.POOL_MAX)
                xs.append(full)
            else:
                full = self._conv(b, h)
                xs.append(full)
                h = full if last else self._pool(full, ops.POOL_MAX)
        h = xs[-1]
        for i, up in enumerate(self.up):
            skip = xs[self.levels - 2 - i]
            if fp32:
                u = ops.convt2x2_f32(h, up["w"], up["b"], out_hw=(skip.shape[2], skip.shape[3]))
            else:
                # any H x W: MaxPool2d floors odd extents, so the upsampled map can be one row / column short of the skip
                # connection; F.pad (unet_parts.py:58-62) fills the high side with zeros — the transposed convolution
                # writes straight into a zeroed tensor of the skip's size
                u = ops.conv_gemm_bf16(h, up["wpack"], up["b"], up["cout"], 4, False, out_hw=(skip.shape[1], skip.shape[2]))
            a, b = up["conv"]
            h = self._conv(b, self._conv(a, skip, x1=u))   # cat([skip, up]) as a two-source K loop
        return h

    def outc(self, feat_nchw_f32: torch.Tensor) -> torch.Tensor:
        out = ops.conv1x1_f32(feat_nchw_f32, self.outc_w, self.outc_b)
        return out

    def gaussian(self, which: str, x: torch.Tensor, segm: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mu, log_sigma) fp32 [B,L] of the prior (segm None) or posterior net."""
        e = self.enc[which]
        fp32 = self.precision == "fp32"
        h = x
        nlev = len(e["layers"])
        for i, (a, b) in enumerate(e["layers"]):
            if i == 0 and segm is not None:
                h = self._conv(a, x, x1=segm) if fp32 else self._conv(a, x, first_x1=segm)
            else:
                h = self._conv(a, h)
            if i < nlev - 1:
                # AvgPool2d(2,2,ceil_mode) in front of the next level: fused into this conv's epilogue
                if not fp32 and b.tc and ops.fused_pool_ok(h.shape[1], h.shape[2]):
                    _, h = ops.conv_gemm_pool_bf16(h, b.wpack, b.b, b.cout, True, ops.POOL_AVG_CEIL, want_full=False)
                else:
                    h = self._pool(self._conv(b, h), ops.POOL_AVG_CEIL)
            else:
                h = self._conv(b, h)
        if fp32:
            return ops.gauss_head_f32(h, e["head_w"], e["head_b"], e["L"])
        return ops.gauss_head_bf16(h, e["head_w"], e["head_b"], e["L"])

    def features_nchw_f32(self, feat: torch.Tensor) -> torch.Tensor:
        return feat if self.precision == "fp32" else ops.nhwc_bf16_to_nchw_f32(feat)

    def fcomb_logits(self, feat_nchw_f32: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """feat fp32 NCHW, z [B,L] or [B,N,L] -> logits [B,C,H,W] or [B,N,C,H,W]."""
        squeeze = z.dim() == 2
        zz = z[:, None, :].contiguous() if squeeze else z.contiguous()
        logits, _ = ops.fcomb_f32(feat_nchw_f32, zz, self.fcomb, want_logits=True, want_sums=False)
        return logits[:, 0] if squeeze else logits

    def fcomb_sums(self, feat: torch.Tensor, mu, sigma, eps, out=None) -> torch.Tensor:
        """Fused N-sample fcomb + softmax + (sum, sum^2): slice_sums [B,2,C,H,W]."""
        if self.h16 is not None and self.fcomb["F"] == 64 and self.fcomb["nl"] <= 6:
            return ops.fcomb_softmax_accum_bf16(feat, mu, sigma, eps, self.fcomb, out=out)
        # fp32 mode, or a head the tensor-core kernel does not cover (F != 64, no_convs_fcomb > 6): CUDA-core kernel
        f = self.features_nchw_f32(feat)
        z = (mu[:, None, :] + sigma[:, None, :] * eps).contiguous()
        _, sums = ops.fcomb_f32(f, z, self.fcomb, want_logits=False, want_sums=True, sums_out=out)
        return sums
