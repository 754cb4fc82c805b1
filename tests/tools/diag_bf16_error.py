#!/usr/bin/env python
"""Where the bf16 mode's probability error comes from (config-2 / config-3 sized slices, random-init trainer model):
features, prior mu / sigma, and probabilities with fp32 vs bf16 inputs to each stage, against the fp32 oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pmu_oracle as O  # noqa: E402
import pmu_b200  # noqa: E402
from pmu_b200.engine import PackedNet  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = 8
sd = O.make_state_dict(seed=0)
vol, _ = O.phantom(D, seed=1234)
eps = torch.randn(2, N, 6, generator=torch.Generator().manual_seed(1))
for plane, s0 in [(0, 40), (1, 77), (2, D - 2)]:
    x = torch.from_numpy(O.plane_slices(vol, plane, s0, 2))
    feat = O.unet_features(sd, x)
    mu, ls = O.gaussian_head(sd, "prior", x)
    sig = torch.exp(ls)
    pk = PackedNet(sd, torch.device("cuda"), "bf16")
    f16 = pk.features_nchw_f32(pk.unet_features(x.cuda())).cpu()
    mu16, ls16 = [t.cpu() for t in pk.gaussian("prior", x.cuda())]
    print(f"plane {plane}: feat max {float(feat.abs().max()):.3f} err max {float((f16 - feat).abs().max()):.4f} "
          f"rms {float((f16 - feat).pow(2).mean().sqrt()):.5f} (feat rms {float(feat.pow(2).mean().sqrt()):.3f}); "
          f"mu err {float((mu16 - mu).abs().max()):.5f} (|mu| {float(mu.abs().max()):.3f}) "
          f"log_sigma err {float((ls16 - ls).abs().max()):.5f} (sigma max {float(sig.max()):.3f})")

    def probs(f, m, s):
        acc = 0
        for n in range(N):
            acc = acc + torch.softmax(O.fcomb(sd, f, m + s * eps[:, n]), 1)
        return acc / N

    ref = probs(feat, mu, sig)
    for name, p in [("bf16 features, fp32 mu/sigma", probs(f16, mu, sig)),
                    ("fp32 features, bf16 mu/sigma", probs(feat, mu16, torch.exp(ls16))),
                    ("both bf16, fp32 fcomb", probs(f16, mu16, torch.exp(ls16)))]:
        e = (p - ref).abs()
        print(f"    {name}: max {float(e.max()):.5f}  p99.9 {float(e.flatten().quantile(0.999)):.5f}  mean {float(e.mean()):.6f}")
    sums = pk.fcomb_sums(pk.unet_features(x.cuda()), mu16.cuda(), torch.exp(ls16).cuda(), eps.cuda().contiguous())
    e = (sums[:, 0].cpu() / N - ref).abs()
    print(f"    full bf16 path (tcgen05 fcomb): max {float(e.max()):.5f}  p99.9 {float(e.flatten().quantile(0.999)):.5f}  mean {float(e.mean()):.6f}")
    logits = O.fcomb(sd, feat, mu + sig * eps[:, 0])
    print(f"    logits range {float(logits.min()):.2f} .. {float(logits.max()):.2f}")
