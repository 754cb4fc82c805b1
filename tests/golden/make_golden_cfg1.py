#!/usr/bin/env python
"""BASELINE config 1 from the REAL reference: Probabilistic U-Net 2D forward + ELBO on ONE synthetic
batch of 4 x 1 x 128 x 128 slices, latent_dim = 6, trainer architecture, CPU.

    python -B tests/golden/make_golden_cfg1.py          (build container only: needs /root/reference)

Pinned inputs: x = randn(4,1,128,128) seed 7, mask = randint(0,3,(4,1,128,128)) seed 8, eps_q = randn(4,6) seed 9
(z_q = mu_q + sigma_q * eps_q, injected by replacing the posterior's rsample for the one elbo() call).  Weights:
oracle.make_state_dict(seed=0), strict-loaded into the reference's own ProbabilisticUnet.  Stored, for eval-mode and
train-mode (batch statistics) BatchNorm: prior / posterior mu and sigma, per-item analytic KL, self.kl,
self.reconstruction_loss, the elbo() value, dice_coeff of the argmax one-hot against (mask == k) for k = 1, 2
(eval.py:42-49) and the reconstruction's probabilities (float16 copy: the comparison budget is 1e-4 / 2e-2 plus the
half-precision storage step).  -> tests/golden/golden_cfg1.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ProbabilisticUnet, UNet, dice_coeff = import_reference()
    sys.path.insert(0, ROOT)
    from oracle import pmu_oracle as O
    sd = O.make_state_dict(seed=0)
    torch.manual_seed(0)
    net = ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[64, 128, 256, 512, 1024],
                            latent_dim=6, no_convs_fcomb=4, beta=10)
    net.load_state_dict(sd, strict=True)
    x = torch.randn(4, 1, 128, 128, generator=torch.Generator().manual_seed(7))
    mask = torch.randint(0, 3, (4, 1, 128, 128), generator=torch.Generator().manual_seed(8)).float()
    eps_q = torch.randn(4, 6, generator=torch.Generator().manual_seed(9))
    out = {"x_seed": np.array(7), "mask_seed": np.array(8), "eps_seed": np.array(9), "eps_q": eps_q.numpy()}
    for tag in ("eval", "train"):
        net.load_state_dict(sd, strict=True)          # train() mutates the running statistics: restore
        net.eval() if tag == "eval" else net.train()
        with torch.no_grad():
            net.forward(x, mask, training=True)
            post = net.posterior_latent_space
            mu_q, sg_q = post.base_dist.loc.clone(), post.base_dist.scale.clone()
            z_q = mu_q + sg_q * eps_q
            post.rsample = lambda *a, **k: z_q          # the one stochastic call of elbo()
            e = net.elbo(mask)
            out[f"{tag}/mu_p"] = net.prior_latent_space.base_dist.loc.numpy()
            out[f"{tag}/sigma_p"] = net.prior_latent_space.base_dist.scale.numpy()
            out[f"{tag}/mu_q"], out[f"{tag}/sigma_q"] = mu_q.numpy(), sg_q.numpy()
            out[f"{tag}/kl_items"] = net.kl_divergence(analytic=True).numpy()
            out[f"{tag}/kl"] = np.array(float(net.kl))
            out[f"{tag}/reconstruction_loss"] = np.array(float(net.reconstruction_loss))
            out[f"{tag}/elbo"] = np.array(float(e))
            prob = torch.softmax(net.reconstruction, 1)
            out[f"{tag}/prob_f16"] = prob.numpy().astype(np.float16)
            lab = torch.argmax(net.reconstruction, 1)
            out[f"{tag}/dice"] = np.array([float(dice_coeff((lab == k).float(), (mask[:, 0] == k).float())) for k in (1, 2)])
            out[f"{tag}/feat_absmean"] = np.array(float(net.unet_features.abs().mean()))
        print(tag, {k.split("/")[1]: (v.tolist() if v.size < 9 else v.shape) for k, v in out.items() if k.startswith(tag)})
    np.savez_compressed(os.path.join(HERE, "golden_cfg1.npz"), **out)
    print("golden_cfg1.npz", os.path.getsize(os.path.join(HERE, "golden_cfg1.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
