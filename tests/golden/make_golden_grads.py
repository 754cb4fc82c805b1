#!/usr/bin/env python
"""Golden GRADIENTS of one training step, produced by the REAL reference model under autograd.

Run in the build container only (needs /root/reference):

    python -B tests/golden/make_golden_grads.py

The reference ProbabilisticUnet (model/probabilistic_unet/probabilistic_unet.py, imported with the
shims of make_golden.py) is put in train() mode and stepped exactly like train.py:85-97 /
probunet_trainer.py:27-39 do: forward(x, segm, training=True), loss = -elbo(segm), loss.backward().
The posterior noise is the first generator draw after torch.manual_seed(SEED) — what
Normal.rsample() consumes inside elbo() — and is stored so the CUDA path / oracle can inject it.

Fixture golden_grads.npz: state_dict before the step (sd/...), inputs, eps_q, elbo / kl / rec,
every parameter gradient (grad/...), and the BatchNorm running statistics after the step (post/...).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, randomise_bn  # noqa: E402

SEED = 9


def main():
    ProbabilisticUnet, _, _ = import_reference()
    torch.manual_seed(0)
    net = ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[4, 8, 16, 32, 64],
                            latent_dim=6, no_convs_fcomb=4, beta=10)
    randomise_bn(net, seed=1)
    out = {}
    for k, v in net.state_dict().items():
        out["sd/" + k] = v.detach().numpy().copy()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(3, 1, 32, 48, generator=g)
    segm = torch.randint(0, 3, (3, 1, 32, 48), generator=g).float()
    out["x"], out["segm"] = x.numpy(), segm.numpy()
    net.train()
    net.forward(x, segm, training=True)
    torch.manual_seed(SEED)
    eps = torch.distributions.utils._standard_normal(net.posterior_latent_space.base_dist.loc.shape,
                                                     dtype=torch.float32, device=x.device)
    torch.manual_seed(SEED)
    elbo = net.elbo(segm)
    z_q = net.posterior_latent_space.base_dist.loc + eps * net.posterior_latent_space.base_dist.scale
    loss = -elbo
    loss.backward()
    out["eps_q"] = eps.numpy()
    out["z_q"] = z_q.detach().numpy()
    out["elbo"] = np.array(float(elbo))
    out["kl"] = np.array(float(net.kl))
    out["rec"] = np.array(float(net.reconstruction_loss))
    n_none = 0
    for k, p in net.named_parameters():
        if p.grad is None:
            n_none += 1
            continue
        out["grad/" + k] = p.grad.detach().numpy().copy()
    for k, v in net.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["post/" + k] = v.detach().numpy().copy()
    # self-check: the stored eps reproduces the reconstruction the reference used
    with torch.no_grad():
        rec_logits = net.fcomb.forward(net.unet_features, z_q)
    assert torch.allclose(rec_logits, net.reconstruction, atol=1e-6), "eps replay does not match rsample()"
    path = os.path.join(HERE, "golden_grads.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB; params without grad: {n_none}")


if __name__ == "__main__":
    main()
