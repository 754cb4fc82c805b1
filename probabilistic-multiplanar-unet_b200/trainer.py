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
        # probunet_trainer.py:25 — an attribute the reference's ProbUNetTrainer never calls (its loss is -elbo); kept so
        # code that touches trainer.criterion keeps working.  It is a plain torch module, not part of any pmu_b200 path.
        self.criterion = torch.nn.BCELoss() if self.net.n_classes == 1 else torch.nn.CrossEntropyLoss()

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

    def mask_to_image(self, masks, prediction=False):
        """probunet_trainer.py:62-92: class indices -> RGB [B,3,H,W] (0 black, 1 blue, 2 green, 3 red); predictions are
        arg-maxed first.  One table lookup on the masks' device instead of the reference's per-pixel Python loop."""
        if self.net.n_classes == 1:
            return (masks >= 0.5).float() if prediction else masks
        colors = torch.tensor([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [1., 0., 0.]], device=masks.device)
        idx = torch.argmax(masks, dim=1) if prediction else masks.squeeze(1).long()
        return colors[idx].permute(0, 3, 1, 2)

    def latent_grid(self, imgs, true_masks, n_preds=3, sigma_scale=1.0, axes=(0, 1)):
        """visualize_sampling.py:11-31: the n_preds x n_preds grid of predictions around the prior mean — one network
        pass and one fcomb launch (ProbabilisticUnet.sample_grid).  Returns (images [B,G,G,3,H,W], logits, z)."""
        with torch.no_grad():
            self.net.forward(imgs, true_masks, training=False)
            logits, z = self.net.sample_grid(n_preds, sigma_scale, axes)
        B, G = logits.shape[0], logits.shape[1]
        img = self.mask_to_image(logits.reshape(B * G * G, *logits.shape[3:]), prediction=True)
        return img.reshape(B, G, G, *img.shape[1:]), logits, z
