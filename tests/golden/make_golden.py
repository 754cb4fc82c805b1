#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python -B tests/golden/make_golden.py

It imports the reference's own ``model/`` and ``dice_loss.py`` with the two shims
SURVEY.md Appendix C describes (a stub ``matplotlib.pyplot`` and torch/nn/np
injected into builtins, because probabilistic_unet.py uses them un-imported),
runs them on seeded inputs and stores inputs + outputs as small .npz fixtures.
``tests/test_oracle_golden.py`` then pins ``oracle/pmu_oracle.py`` against them.

Fixtures:
  golden_small.npz    reference-constructed ProbabilisticUnet([4,8,16,32,64]) with
                      randomised BN statistics: state_dict + eval-mode and
                      train-mode forward / fcomb / sample_at / kl / elbo outputs.
  golden_trainer.npz  the trainer model ([64..1024], probunet_trainer.py:16) loaded
                      (strict=True) with oracle.make_state_dict(seed=0) weights — only
                      the outputs are stored; the test regenerates the weights by seed.
  golden_dice.npz     dice_coeff on seeded tensors.
"""
import builtins
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMU_REFERENCE", "/root/reference/Probabilistic-Multiplanar-Unet")


def import_reference():
    shim = tempfile.mkdtemp(prefix="pmu_shim_")
    os.makedirs(os.path.join(shim, "matplotlib"))
    open(os.path.join(shim, "matplotlib", "__init__.py"), "w").close()
    open(os.path.join(shim, "matplotlib", "pyplot.py"), "w").close()
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    builtins.torch, builtins.nn, builtins.np = torch, nn, np
    from model import ProbabilisticUnet, UNet  # noqa
    from dice_loss import dice_coeff  # noqa
    return ProbabilisticUnet, UNet, dice_coeff


def randomise_bn(net, seed=1):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
            m.running_var.copy_(0.5 + torch.rand(c, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(c, generator=g))
            m.bias.data.copy_(torch.randn(c, generator=g) * 0.1)


def run_model(net, x, segm, z, z1, z_q_seed, out, tag):
    """forward(training=True) then everything the hot path calls; tag = 'eval' | 'train'."""
    with torch.no_grad():
        net.forward(x, segm, training=True)
        out[f"{tag}/features"] = net.unet_features.numpy()
        out[f"{tag}/mu_p"] = net.prior_latent_space.base_dist.loc.numpy()
        out[f"{tag}/sigma_p"] = net.prior_latent_space.base_dist.scale.numpy()
        out[f"{tag}/mu_q"] = net.posterior_latent_space.base_dist.loc.numpy()
        out[f"{tag}/sigma_q"] = net.posterior_latent_space.base_dist.scale.numpy()
        out[f"{tag}/fcomb_logits"] = net.fcomb.forward(net.unet_features, z).numpy()
        out[f"{tag}/kl"] = net.kl_divergence(analytic=True).numpy()
        # elbo with a known posterior sample: replay the generator
        torch.manual_seed(z_q_seed)
        z_q = net.posterior_latent_space.rsample()
        torch.manual_seed(z_q_seed)
        e = net.elbo(segm)
        out[f"{tag}/z_q"] = z_q.numpy()
        out[f"{tag}/elbo"] = np.array(float(e))
        out[f"{tag}/elbo_kl"] = np.array(float(net.kl))
        out[f"{tag}/elbo_rec"] = np.array(float(net.reconstruction_loss))
        out[f"{tag}/elbo_logits"] = net.reconstruction.numpy()
        # reconstruct(z_posterior=) hook
        out[f"{tag}/reconstruct"] = net.reconstruct(z_posterior=z_q).numpy()
        if x.shape[0] == 1:
            out[f"{tag}/sample_at"] = net.sample_at(z1).numpy()


def main():
    ProbabilisticUnet, UNet, dice_coeff = import_reference()
    sys.path.insert(0, ROOT)
    from oracle import pmu_oracle as O

    # ---------------- small model, reference-constructed weights ----------------
    torch.manual_seed(0)
    net = ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[4, 8, 16, 32, 64],
                            latent_dim=6, no_convs_fcomb=4, beta=10)
    randomise_bn(net, seed=1)
    out = {}
    for k, v in net.state_dict().items():
        out["sd/" + k] = v.detach().numpy().copy()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 1, 32, 48, generator=g)
    segm = torch.randint(0, 3, (2, 1, 32, 48), generator=g).float()
    z = torch.randn(2, 6, generator=g)
    out["x"], out["segm"], out["z"] = x.numpy(), segm.numpy(), z.numpy()
    net.eval()
    run_model(net, x, segm, z, None, 9, out, "eval")
    # batch-1 run for sample_at (reference only supports batch 1, probabilistic_unet.py:247)
    out1 = {}
    run_model(net, x[:1], segm[:1], z[:1], z[0], 9, out1, "eval")
    out["eval/sample_at_b1"] = out1["eval/sample_at"]
    # plain UNet with apply_last_layer=True (unet_model.py:31-54)
    with torch.no_grad():
        net.unet.apply_last_layer = True
        out["eval/unet_out"] = net.unet.forward(x).numpy()
        net.unet.apply_last_layer = False
    # train-mode BN (config 1 semantics: forward(training=True)+elbo with batch statistics).
    # NOTE train() mutates running stats; the state_dict above was captured before and
    # nothing eval-mode is computed after this point.
    net.train()
    run_model(net, x, segm, z, None, 9, out, "train")
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("golden_small.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")

    # ---------------- trainer model, oracle-seeded weights ----------------
    sd = O.make_state_dict(seed=0)
    torch.manual_seed(0)
    big = ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[64, 128, 256, 512, 1024],
                            latent_dim=6, no_convs_fcomb=4, beta=10)
    missing = big.load_state_dict(sd, strict=True)   # validates the key schema of make_state_dict
    print("load_state_dict(strict=True):", missing)
    big.eval()
    g = torch.Generator().manual_seed(11)
    xb = torch.rand(2, 1, 32, 32, generator=g)
    sb = torch.randint(0, 3, (2, 1, 32, 32), generator=g).float()
    zb = torch.randn(2, 6, generator=g)
    ob = {"x": xb.numpy(), "segm": sb.numpy(), "z": zb.numpy()}
    with torch.no_grad():
        big.forward(xb, sb, training=True)
        ob["features"] = big.unet_features.numpy()
        ob["mu_p"] = big.prior_latent_space.base_dist.loc.numpy()
        ob["sigma_p"] = big.prior_latent_space.base_dist.scale.numpy()
        ob["mu_q"] = big.posterior_latent_space.base_dist.loc.numpy()
        ob["sigma_q"] = big.posterior_latent_space.base_dist.scale.numpy()
        ob["fcomb_logits"] = big.fcomb.forward(big.unet_features, zb).numpy()
        ob["kl"] = big.kl_divergence(analytic=True).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_trainer.npz"), **ob)
    print("golden_trainer.npz", sum(v.nbytes for v in ob.values()) / 1e6, "MB raw")

    # ---------------- dice ----------------
    g = torch.Generator().manual_seed(21)
    p = (torch.rand(3, 16, 16, generator=g) > 0.5).float()
    t = (torch.rand(3, 16, 16, generator=g) > 0.5).float()
    zp = torch.zeros(2, 8, 8)
    np.savez_compressed(os.path.join(HERE, "golden_dice.npz"), p=p.numpy(), t=t.numpy(),
                        d=np.array(float(dice_coeff(p, t))), d_empty=np.array(float(dice_coeff(zp, zp))),
                        d_soft=np.array(float(dice_coeff(torch.rand(3, 16, 16, generator=torch.Generator().manual_seed(22)), t))))
    print("golden_dice.npz written")


if __name__ == "__main__":
    main()
