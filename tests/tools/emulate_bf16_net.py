#!/usr/bin/env python
"""CPU emulation of the 16-bit tensor-core mode's numerics, layer by layer: which roundings make the worst pixel of a
single view's N-sample mean probability (tests/test_gpu_model.py::_spot_check_planes) and what each mitigation buys.

The emulation folds eval-mode BatchNorm like engine.PackedNet, rounds the folded weights and every stored activation
with a configurable quantiser (bf16 / f16 / none) and accumulates in fp32 (what tcgen05 kind::f16 does: exact
products of 16-bit operands, fp32 accumulation).

    python tests/tools/emulate_bf16_net.py [D] [plane] [slice]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pmu_oracle as O  # noqa: E402

BN_EPS = 1e-5


def q_bf16(t):
    return t.to(torch.bfloat16).float()


def q_f16(t):
    return t.to(torch.float16).float()


def q_none(t):
    return t


def fold(sd, conv, bn):
    w, b = sd[conv + ".weight"], sd[conv + ".bias"]
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
    return w * s[:, None, None, None], (b - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]


def net(sd, x, mu, sigma, eps, qw, qa, qw_last=None, qa_last=None, qfw=None, qfa=None, first_exact=True):
    """qw / qa: weight / activation quantiser of the U-Net; *_last: of the last decoder DoubleConv (up4) and its
    output features; qfw / qfa: of the fcomb weights / hidden activations.  Returns mean probability [C,H,W]."""
    qw_last, qa_last = qw_last or qw, qa_last or qa
    qfw, qfa = qfw or qw, qfa or qa

    def dconv(h, p, w_q, a_q, first=False):
        for c, bn in ((".double_conv.0", ".double_conv.1"), (".double_conv.3", ".double_conv.4")):
            w, b = fold(sd, p + c, p + bn)
            if first and first_exact:         # the 1 -> 64 layer splits its fp32 input into hi + lo: ~fp32 input, bf16 weights
                h = a_q(F.relu(F.conv2d(h, w_q(w), b, padding=1)))
                first = False
            else:
                h = a_q(F.relu(F.conv2d(h, w_q(w), b, padding=1)))
        return h

    L = O.num_levels(sd)
    xs = [dconv(x, "unet.inc", qw, qa, first=True)]
    for i in range(L - 1):
        xs.append(dconv(F.max_pool2d(xs[i], 2), f"unet.down_blocks.{i}.maxpool_conv.1", qw, qa))
    h = xs[-1]
    for i in range(L - 1):
        last = i == L - 2
        wq, aq = (qw_last, qa_last) if last else (qw, qa)
        up = aq(F.conv_transpose2d(h, wq(sd[f"unet.up_blocks.{i}.up.weight"]), sd[f"unet.up_blocks.{i}.up.bias"], stride=2))
        h = dconv(torch.cat([xs[L - 2 - i], up], 1), f"unet.up_blocks.{i}.conv", wq, aq)
    feat = h
    # fcomb: layer 0 split into the feature GEMM (fp32 accumulate) + exact fp32 per-sample bias, hidden activations stored qfa
    w0 = sd["fcomb.layers.0.weight"][:, :, 0, 0]
    Fch = feat.shape[1]
    G = F.conv2d(feat, qfw(w0[:, :Fch])[:, :, None, None])
    acc = 0
    N = eps.shape[0]
    for n in range(N):
        z = mu + sigma * eps[n]
        zb = w0[:, Fch:] @ z + sd["fcomb.layers.0.bias"]
        hcur = qfa(F.relu(G + zb[None, :, None, None]))
        i = 1
        while f"fcomb.layers.{2 * i}.weight" in sd:
            hcur = qfa(F.relu(F.conv2d(hcur, qfw(sd[f"fcomb.layers.{2 * i}.weight"]), sd[f"fcomb.layers.{2 * i}.bias"])))
            i += 1
        logits = F.conv2d(hcur, qfw(sd["fcomb.last_layer.weight"]), sd["fcomb.last_layer.bias"])
        acc = acc + torch.softmax(logits, 1)
    return (acc / N)[0]


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    plane = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    s = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    N = 16
    torch.set_num_threads(os.cpu_count())
    sd = O.make_state_dict(seed=0)
    vol, _ = O.phantom(D, seed=1234)
    eps_all = torch.randn(3, D, N, 6, generator=torch.Generator().manual_seed(4321))
    x = torch.from_numpy(O.plane_slices(vol, plane, s, 1))
    with torch.no_grad():
        mu, ls = O.gaussian_head(sd, "prior", x)
        mu, sigma, eps = mu[0], torch.exp(ls[0]), eps_all[plane, s]
        ref = net(sd, x, mu, sigma, eps, q_none, q_none)
        rows = [
            ("all bf16 (today)", dict(qw=q_bf16, qa=q_bf16)),
            ("U-Net bf16, fcomb fp32", dict(qw=q_bf16, qa=q_bf16, qfw=q_none, qfa=q_none)),
            ("U-Net fp32, fcomb bf16", dict(qw=q_none, qa=q_none, qfw=q_bf16, qfa=q_bf16)),
            ("U-Net bf16, fcomb f16", dict(qw=q_bf16, qa=q_bf16, qfw=q_f16, qfa=q_f16)),
            ("U-Net bf16 weights only (fp32 activations), fcomb fp32", dict(qw=q_bf16, qa=q_none, qfw=q_none, qfa=q_none)),
            ("U-Net bf16 activations only (fp32 weights), fcomb fp32", dict(qw=q_none, qa=q_bf16, qfw=q_none, qfa=q_none)),
            ("U-Net bf16 but last DoubleConv + up4.T f16, fcomb f16", dict(qw=q_bf16, qa=q_bf16, qw_last=q_f16, qa_last=q_f16, qfw=q_f16, qfa=q_f16)),
            ("all f16", dict(qw=q_f16, qa=q_f16)),
        ]
        print(f"{D}^2 slice (plane {plane}, s {s}), N = {N}: |mean prob - fp32| per pixel")
        for name, kw in rows:
            e = (net(sd, x, mu, sigma, eps, **kw) - ref).abs()
            print(f"  {name:62s} max {float(e.max()):.4f}  p99.9 {float(e.flatten().quantile(0.999)):.4f}  mean {float(e.mean()):.5f}")


if __name__ == "__main__":
    main()
