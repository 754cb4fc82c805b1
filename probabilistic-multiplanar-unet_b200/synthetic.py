"""Seeded synthetic inputs for bench.py / smoke(): a phantom volume and random-init weights of the
reference's trainer architecture (there is no network for datasets or checkpoints).  Host-side
setup only — nothing here is on the timed path."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .model import ProbabilisticUnet

TRAINER_FILTERS = [64, 128, 256, 512, 1024]     # trainer/probunet_trainer.py:16


def phantom_volume(D: int, seed: int = 1234) -> torch.Tensor:
    """0.6 * two nested offset ellipsoids + 0.4 * uniform noise, fp32 [D,D,D] in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    ax = torch.linspace(-1, 1, D)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    outer = ((X - 0.05) / 0.80) ** 2 + ((Y + 0.10) / 0.65) ** 2 + (Z / 0.70) ** 2 <= 1.0
    inner = ((X - 0.15) / 0.40) ** 2 + ((Y + 0.05) / 0.30) ** 2 + ((Z - 0.10) / 0.35) ** 2 <= 1.0
    ph = torch.zeros(D, D, D)
    ph[outer] = 1.0
    ph[inner] = 0.5
    return (0.6 * ph + 0.4 * torch.rand(D, D, D, generator=g)).float().contiguous()


def phantom_labels(D: int) -> torch.Tensor:
    """Label volume of the same phantom: 0 background, 1 outer shell, 2 inner ellipsoid; fp32 [D,D,D] (the reference's
    mask_type is float32, probunet_trainer.py:14)."""
    ax = torch.linspace(-1, 1, D)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    outer = ((X - 0.05) / 0.80) ** 2 + ((Y + 0.10) / 0.65) ** 2 + (Z / 0.70) ** 2 <= 1.0
    inner = ((X - 0.15) / 0.40) ** 2 + ((Y + 0.05) / 0.30) ** 2 + ((Z - 0.10) / 0.35) ** 2 <= 1.0
    lab = torch.zeros(D, D, D)
    lab[outer] = 1.0
    lab[inner] = 2.0
    return lab.contiguous()


def trainer_state_dict(seed: int = 0, num_filters: Sequence[int] = TRAINER_FILTERS, num_classes: int = 3,
                       latent_dim: int = 6, no_convs_fcomb: int = 4):
    """Random-init weights of the trainer model (the reference's initialisers) with randomised
    BatchNorm statistics so that BN folding does real work."""
    torch.manual_seed(seed)
    net = ProbabilisticUnet(1, num_classes, list(num_filters), latent_dim, no_convs_fcomb, beta=10)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                c = m.num_features
                m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))
                m.weight.copy_(0.5 + torch.rand(c, generator=g))
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
        for which in (net.prior, net.posterior):
            which.conv_layer.bias.copy_(torch.randn(2 * latent_dim, generator=g) * 0.3)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}
