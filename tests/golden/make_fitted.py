#!/usr/bin/env python
"""Fit a small ProbabilisticUnet([64,128]) to the nested-ellipsoid phantom on the CPU (oracle
functional forward + autograd) and store its weights (fp16-rounded, so every consumer sees
the same values) as tests/golden/fitted_small.npz.

Purpose: a CONFIDENT model for the north-star "per-volume Dice agreement >= 0.999" check
between the bf16/tcgen05 path and the fp32 oracle — random weights give near-tied softmaxes
where argmax agreement is meaningless (SURVEY.md §7 hard parts).

    python -B tests/golden/make_fitted.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pmu_oracle as O  # noqa


def main():
    torch.manual_seed(0)
    sd = O.make_state_dict((64, 128), num_classes=3, latent_dim=6, no_convs_fcomb=4, seed=5)
    params = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k and not k.startswith("posterior")]
    for k in params:
        sd[k] = sd[k].clone().requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in params], lr=2e-3)
    D = 48
    vol, lab = O.phantom(D, seed=99)
    xs, ys = [], []
    for p in range(3):
        xs.append(torch.from_numpy(O.plane_slices(vol, p)))
        ys.append(torch.from_numpy(np.stack([O.sample_slice(lab, p, s) for s in range(D)])).long())
    X, Y = torch.cat(xs), torch.cat(ys)
    g = torch.Generator().manual_seed(1)
    for step in range(400):
        idx = torch.randint(0, X.shape[0], (16,), generator=g)
        x, y = X[idx], Y[idx]
        feat = O.unet_features(sd, x)
        mu, ls = O.gaussian_head(sd, "prior", x)
        z = mu + torch.exp(ls) * torch.randn(mu.shape, generator=g)
        logits = O.fcomb(sd, feat, z)
        loss = F.cross_entropy(logits, y) + 1e-3 * (mu ** 2 + ls ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 50 == 0 or step == 399:
            acc = (logits.argmax(1) == y).float().mean()
            print(f"step {step}: loss {float(loss):.4f} acc {float(acc):.4f}")
    out = {}
    for k, v in sd.items():
        v = v.detach()
        out[k] = v.numpy().astype(np.float16) if v.dtype.is_floating_point else v.numpy()
    np.savez_compressed(os.path.join(HERE, "fitted_small.npz"), **out)
    print("fitted_small.npz", os.path.getsize(os.path.join(HERE, "fitted_small.npz")) / 1e6, "MB")
    # report how decisive the fitted model is on a fresh phantom
    sd32 = {k: torch.from_numpy(v.astype(np.float32)) if v.dtype == np.float16 else torch.from_numpy(v) for k, v in out.items()}
    vol2, lab2 = O.phantom(32, seed=7)
    eps = torch.randn(3, 32, 4, 6, generator=torch.Generator().manual_seed(4321))
    r = O.multiplanar_predict(vol2, sd32, eps, 4, batch=32)
    top2 = torch.topk(r["mean"], 2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    print("voxels with margin < 0.02:", int((margin < 0.02).sum()), "of", margin.numel(),
          "label acc", float((r["mean"].argmax(1) == torch.from_numpy(lab2).long()).float().mean()))


if __name__ == "__main__":
    main()
