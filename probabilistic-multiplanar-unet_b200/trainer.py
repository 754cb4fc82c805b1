"""ProbUNetTrainer mirror (reference trainer/probunet_trainer.py:10-60, trainer/trainer.py): the
adapter eval.py / train.py / visualize_sampling.py call.  predict / loss / eval keep the
reference's argument meaning; the arithmetic runs in the CUDA kernels behind ProbabilisticUnet."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .dice_loss import dice_coeff
from .model import ProbabilisticUnet


class Trainer:
    def predict(self, imgs, masks):
        raise NotImplementedError

    def eval(self, imgs, true_masks, masks_pred):
        raise NotImplementedError

    def loss(self, imgs, true_masks, masks_pred):
        raise NotImplementedError


class ProbUNetTrainer(Trainer):
    def __init__(self, device, n_channels=1, n_classes=1, load_model=None, latent_dim=6, beta=10, precision="fp32"):
        self.device = device
        self.mask_type = torch.float32
        self.name = "probunet"
        self.net = ProbabilisticUnet(input_channels=n_channels, num_classes=n_classes,
                                     num_filters=[64, 128, 256, 512, 1024], latent_dim=latent_dim,
                                     no_convs_fcomb=4, beta=beta)
        if load_model is not None:
            self.net.load_state_dict(torch.load(load_model, map_location=device), strict=False)
        self.net = self.net.to(device)
        self.net.set_precision(precision)

    def predict(self, imgs, true_masks, z=None):
        """probunet_trainer.py:27-32 (train == grad enabled decides posterior + rsample)."""
        train = torch.is_grad_enabled()
        self.net.forward(imgs, true_masks, training=train)
        return self.net.sample(testing=not train) if z is None else self.net.sample_at(z)

    def loss(self, imgs, true_masks, masks_pred):
        return -self.net.elbo(true_masks)

    def eval(self, imgs, true_masks, masks_pred):
        """probunet_trainer.py:41-60: per-class Dice of the argmax one-hot vs (mask == k)."""
        if self.net.n_classes == 1:
            return np.array([dice_coeff((masks_pred > 0.5).float(), true_masks).item()])
        B, C, H, W = masks_pred.shape
        # argmax is invariant under softmax; layout [X=B, C, Y*Z=H*W] matches pmu_argmax_dice_sums
        s = ops.argmax_dice_sums(masks_pred.contiguous().float(), true_masks.contiguous().float().reshape(B, H, W))
        d = (2.0 * s[:, 0] + 1e-6) / (s[:, 1] + s[:, 2] + 1e-6)
        return d.cpu().numpy()
