#!/usr/bin/env python
"""More golden vectors from the REAL reference model for the less-travelled arguments of the drop-in API
(SURVEY.md §8 rows a15-a17): the Monte-Carlo KL estimate, the ELBO built on it, and the posterior-mean reconstruction.

    python -B tests/golden/make_golden_small_extra.py          (build container only: needs /root/reference)

The small reference model of golden_small.npz is rebuilt (reference constructor, weights loaded from the fixture's own
state dict with strict=True, BatchNorm in eval mode) and run on the fixture's inputs:
  kl_divergence(analytic=False, z_posterior=z_q)          probabilistic_unet.py:274-278   log q(z) - log p(z)  [B]
  elbo(segm, analytic_kl=False)                            :281-308 with the posterior draw replayed from a seed
  reconstruct(use_posterior_mean=True), elbo(segm, reconstruct_posterior_mean=True)
                                                           :251-262, :292 — both RAISE in the reference (`.loc` on Independent)
Fixture: golden_small_extra.npz; tests/test_oracle_golden.py::test_small_model_mc_kl_and_posterior_mean.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ProbabilisticUnet, _, _ = import_reference()
    g = np.load(os.path.join(HERE, "golden_small.npz"))
    net = ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[4, 8, 16, 32, 64], latent_dim=6,
                            no_convs_fcomb=4, beta=10)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    print("load_state_dict(strict=True):", net.load_state_dict(sd, strict=True))
    net.eval()
    x, segm, z_q = torch.from_numpy(g["x"]), torch.from_numpy(g["segm"]), torch.from_numpy(g["eval/z_q"])
    out = {}
    with torch.no_grad():
        net.forward(x, segm, training=True)
        assert np.allclose(net.unet_features.numpy(), g["eval/features"], atol=1e-6)      # same model as the fixture
        out["kl_mc"] = net.kl_divergence(analytic=False, z_posterior=z_q).numpy()
        torch.manual_seed(9)                                    # the seed golden_small's eval/z_q was drawn with
        e = net.elbo(segm, analytic_kl=False)
        out["elbo_mc"] = np.float64(float(e))
        out["elbo_mc_kl"] = np.float64(float(net.kl))
        out["elbo_mc_rec"] = np.float64(float(net.reconstruction_loss))
        # reconstruct(use_posterior_mean=True) and elbo(reconstruct_posterior_mean=True) read `.loc` of the Independent
        # wrapper (probabilistic_unet.py:258), which torch.distributions does not forward: the reference RAISES here.
        # The fixture records that; the drop-in reads base_dist.loc instead (documented deviation).
        for key, call in (("reconstruct_mean_raises", lambda: net.reconstruct(use_posterior_mean=True)),
                          ("elbo_posterior_mean_raises", lambda: net.elbo(segm, reconstruct_posterior_mean=True))):
            try:
                call()
                out[key] = np.int64(0)
            except AttributeError as ex:
                out[key] = np.int64(1)
                print(f"{key}: AttributeError: {ex}")
    path = os.path.join(HERE, "golden_small_extra.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}:", {k: (v.shape if hasattr(v, 'shape') and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
