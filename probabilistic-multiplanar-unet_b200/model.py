"""Drop-in model API: ProbabilisticUnet / UNet with the reference's constructor signatures,
method names, side-effect attributes and state_dict key schema — executed by the sm_100a
kernels in libpmu_b200.so instead of torch ops.

Reference interface mirrored (paths relative to Probabilistic-Multiplanar-Unet/):
  ProbabilisticUnet.__init__/forward/sample/sample_at/reconstruct/kl_divergence/elbo
      model/probabilistic_unet/probabilistic_unet.py:194-308
  UNet.__init__/forward                           model/unet/unet_model.py:10-54
  state_dict keys (unet.inc.double_conv.0.weight, unet.up_blocks.0.up.weight,
      prior.encoder.layers.7.weight, fcomb.last_layer.weight, ...)  — reference checkpoints
      load unchanged with load_state_dict(..., strict=False) (trainer/probunet_trainer.py:18-22).

The nn.Module tree below only HOLDS parameters under the reference's names; no torch op of
these containers is ever called.  Execution goes PackedNet -> ops -> C-ABI.  There is no
CPU / eager fallback: tensors must live on a CUDA device and the extension must be built.

Differences from the (partly broken) reference, all listed in DESIGN.md:
  * inference uses eval-mode BatchNorm (folded).  With autograd enabled on a ProbabilisticUnet in
    train() mode, forward()/elbo() run the fp32 TRAINING path (train-mode BatchNorm, explicit
    backward kernels, train_engine.py); combinations that path does not cover (bf16 precision,
    eval-mode BN with gradients, a bare UNet) raise NotImplementedError instead of silently
    falling back to ATen.
  * sample()/elbo()/reconstruct() accept keyword-only z= / eps= to inject latents (the
    reference has no such hook except sample_at); positional signatures are unchanged.
  * sample_at accepts [L] (reference behaviour, batch 1... broadcast) and [B,L].
"""
from __future__ import annotations

import math
import warnings
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.distributions import Independent, Normal

from . import ops
from . import train_engine
from .engine import PackedNet


# ------------------------------------------------------------------------------------
# parameter containers (names/shapes == reference; forward is never used)
# ------------------------------------------------------------------------------------
def _he_(conv: nn.Module):
    """init_weights (model/probabilistic_unet/utils.py:15-20): kaiming-normal weight,
    truncated-normal(std 1e-3) bias."""
    nn.init.kaiming_normal_(conv.weight, mode="fan_in", nonlinearity="relu")
    nn.init.trunc_normal_(conv.bias, mean=0.0, std=0.001, a=-0.002, b=0.002)


def _orth_(conv: nn.Module):
    """init_weights_orthogonal_normal (utils.py:22-26)."""
    nn.init.orthogonal_(conv.weight)
    nn.init.trunc_normal_(conv.bias, mean=0.0, std=0.001, a=-0.002, b=0.002)


class DoubleConv(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(cin, cout))


class Up(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(cin, cout)


class OutConv(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1)


def _check_filters(num_filters: Sequence[int]):
    for a, b in zip(num_filters[:-1], num_filters[1:]):
        if b != 2 * a:
            raise ValueError(
                f"num_filters must double at every level (got {list(num_filters)}): the reference U-Net's Up block "
                f"assumes it (model/unet/unet_parts.py:52-53) and crashes otherwise")


def _no_autograd(module: nn.Module, what: str):
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError(
            f"{what}: this call has no backward kernels (the training path covers ProbabilisticUnet.forward(training=True) "
            f"+ elbo() in train() mode, fp32). Call under torch.no_grad() for inference; there is no ATen fallback.")


class _PackedMixin:
    _pack_prefix = ""

    def _state_version(self):
        # (storage, version) per tensor: parameter replacement and load_state_dict(assign=True) change the storage,
        # in-place updates the version.  Writes through `.data` bump neither: call invalidate_pack() after them.
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def invalidate_pack(self):
        """Drop the packed (BatchNorm-folded, 16-bit) weights; the next inference call re-packs from the parameters."""
        object.__setattr__(self, "_pack_key", None)
        return self

    def packed(self) -> PackedNet:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pmu_b200 models run on CUDA only (no CPU fallback); call .to('cuda')")
        key = (dev, self.precision, self._state_version())
        if getattr(self, "_pack_key", None) != key:
            sd = {self._pack_prefix + k: v.detach() for k, v in self.state_dict().items()}
            object.__setattr__(self, "_pack", PackedNet(sd, dev, self.precision))
            object.__setattr__(self, "_pack_key", key)
        return self._pack

    def set_precision(self, precision: str):
        if precision not in ("fp32", "f16", "bf16"):
            raise ValueError("precision must be 'fp32', 'f16' (tensor-core inference format; training runs its GEMMs "
                             "in bf16) or 'bf16'")
        self.precision = precision
        return self


class UNet(nn.Module, _PackedMixin):
    """model/unet/unet_model.py:9-54."""
    _pack_prefix = "unet."

    def __init__(self, n_channels, n_classes, num_filters=[64, 128, 256, 512, 1024], bilinear=False,
                 apply_last_layer=True):
        super().__init__()
        if bilinear:
            raise NotImplementedError("bilinear=True is never used by the reference's probabilistic path")
        _check_filters(num_filters)
        self.n_channels, self.n_classes = n_channels, n_classes
        self.bilinear, self.apply_last_layer = bilinear, apply_last_layer
        self.num_filters = num_filters
        self.precision = "fp32"
        self.down_blocks = nn.ModuleList()
        ups = []
        self.inc = DoubleConv(n_channels, num_filters[0])
        self.outc = OutConv(num_filters[0], n_classes)
        for i in range(len(num_filters) - 1):
            self.down_blocks.append(Down(num_filters[i], num_filters[i + 1]))
            ups.append(Up(num_filters[i + 1], num_filters[i]))
        self.up_blocks = nn.ModuleList(ups[::-1])   # deepest first (unet_model.py:29)

    def forward(self, x):
        _no_autograd(self, "UNet.forward")
        pk = self.packed()
        feat = pk.features_nchw_f32(pk.unet_features(x.contiguous().float()))
        if not self.apply_last_layer:
            return feat
        out = pk.outc(feat)
        # torch.sigmoid on [B,1,H,W] for the single-class case (unet_model.py:48-49) — not on the
        # probabilistic path (apply_last_layer=False there).
        return torch.sigmoid(out) if self.n_classes == 1 else out


class Encoder(nn.Module):
    """probabilistic_unet.py:11-53 (parameter container)."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, posterior=False):
        super().__init__()
        self.input_channels = input_channels + (1 if posterior else 0)
        self.num_filters = num_filters
        layers: List[nn.Module] = []
        out_dim = None
        for i, nf in enumerate(num_filters):
            in_dim = self.input_channels if i == 0 else out_dim
            out_dim = nf
            if i != 0:
                layers.append(nn.AvgPool2d(kernel_size=2, stride=2, padding=0, ceil_mode=True))
            layers += [nn.Conv2d(in_dim, out_dim, 3, padding=1), nn.BatchNorm2d(out_dim), nn.ReLU(inplace=True)]
            for _ in range(no_convs_per_block - 1):
                layers += [nn.Conv2d(out_dim, out_dim, 3, padding=1), nn.BatchNorm2d(out_dim), nn.ReLU(inplace=True)]
        self.layers = nn.Sequential(*layers)
        for m in self.layers:
            if isinstance(m, nn.Conv2d):
                _he_(m)


class AxisAlignedConvGaussian(nn.Module):
    """probabilistic_unet.py:55-114 (parameter container)."""

    def __init__(self, input_channels, num_filters, no_convs_per_block, latent_dim, posterior=False):
        super().__init__()
        self.input_channels, self.num_filters = input_channels, num_filters
        self.no_convs_per_block, self.latent_dim, self.posterior = no_convs_per_block, latent_dim, posterior
        self.name = "Posterior" if posterior else "Prior"
        self.encoder = Encoder(input_channels, num_filters, no_convs_per_block, posterior=posterior)
        self.conv_layer = nn.Conv2d(num_filters[-1], 2 * latent_dim, (1, 1), stride=1)
        nn.init.kaiming_normal_(self.conv_layer.weight, mode="fan_in", nonlinearity="relu")
        nn.init.normal_(self.conv_layer.bias)


class Fcomb(nn.Module):
    """probabilistic_unet.py:116-181 (parameter container)."""

    def __init__(self, num_filters, latent_dim, num_output_channels, num_classes, no_convs_fcomb):
        super().__init__()
        self.num_channels, self.num_classes = num_output_channels, num_classes
        self.num_filters, self.latent_dim, self.no_convs_fcomb = num_filters, latent_dim, no_convs_fcomb
        self.name = "Fcomb"
        layers: List[nn.Module] = [nn.Conv2d(num_filters[0] + latent_dim, num_filters[0], kernel_size=1),
                                   nn.ReLU(inplace=True)]
        for _ in range(no_convs_fcomb - 2):
            layers += [nn.Conv2d(num_filters[0], num_filters[0], kernel_size=1), nn.ReLU(inplace=True)]
        self.layers = nn.Sequential(*layers)
        self.last_layer = nn.Conv2d(num_filters[0], num_classes, kernel_size=1)
        for m in list(self.layers) + [self.last_layer]:
            if isinstance(m, nn.Conv2d):
                _orth_(m)


def latent_grid_z(mu: torch.Tensor, sigma: torch.Tensor, n_preds: int = 3, sigma_scale: float = 1.0, axes=(0, 1)) -> torch.Tensor:
    """The z list of visualize_sampling.py:21-26 for every slice: mu, sigma [B, L] -> z [B, G, G, L] with
    z[b, i, j] = mu[b] except z[axes[0]] = s_i * sigma_scale * sigma[axes[0]] + mu[axes[0]] and z[axes[1]] likewise with
    s_j; s = -(n_preds // 2) .. n_preds // 2.  (The reference scales sigma by 40 first, visualize_sampling.py:78, and
    walks z_0 over rows, z_1 over columns.)  Pure tensor code on mu's device."""
    steps = torch.arange(-(n_preds // 2), n_preds // 2 + 1, device=mu.device, dtype=torch.float32)
    G = steps.numel()
    z = mu[:, None, None, :].repeat(1, G, G, 1)
    a0, a1 = axes
    z[:, :, :, a0] = steps[None, :, None] * sigma_scale * sigma[:, None, None, a0] + mu[:, None, None, a0]
    z[:, :, :, a1] = steps[None, None, :] * sigma_scale * sigma[:, None, None, a1] + mu[:, None, None, a1]
    return z


class ProbabilisticUnet(nn.Module, _PackedMixin):
    """Drop-in for model/probabilistic_unet/probabilistic_unet.py:184-308."""

    def __init__(self, input_channels=1, num_classes=1, num_filters=[32, 64, 128, 192], latent_dim=6,
                 no_convs_fcomb=3, beta=1.0):
        super().__init__()
        _check_filters(num_filters)
        if no_convs_fcomb < 2:
            raise ValueError("no_convs_fcomb must be >= 2")
        self.n_channels, self.n_classes = input_channels, num_classes
        self.num_filters, self.latent_dim = num_filters, latent_dim
        self.no_convs_per_block, self.no_convs_fcomb = 2, no_convs_fcomb
        self.initializers = {"w": "he_normal", "b": "normal"}
        self.beta = beta
        self.z_prior_sample = 0
        self.precision = "fp32"
        self.unet = UNet(input_channels, num_classes, num_filters, apply_last_layer=False)
        self.prior = AxisAlignedConvGaussian(input_channels, num_filters, 2, latent_dim)
        self.posterior = AxisAlignedConvGaussian(input_channels, num_filters, 2, latent_dim, posterior=True)
        self.fcomb = Fcomb(num_filters, latent_dim, input_channels, num_classes, no_convs_fcomb)
        self.posterior_latent_space = None
        self.prior_latent_space = None
        self.unet_features = None
        self._warned_train = False

    # -- helpers ------------------------------------------------------------------
    def _dist(self, mu, log_sigma):
        return Independent(Normal(loc=mu, scale=torch.exp(log_sigma)), 1)

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _enter(self, what):
        _no_autograd(self, what)
        if self.training and not self._warned_train:
            warnings.warn("ProbabilisticUnet is in train() mode; the CUDA path always uses eval-mode (folded) "
                          "BatchNorm — call net.eval() (the reference's eval.py forgets to, SURVEY.md App. B #6)")
            self._warned_train = True

    # -- reference API --------------------------------------------------------------
    @property
    def unet_features(self):
        """probabilistic_unet.py:223: the U-Net feature map, fp32 NCHW.  The tensor-core training step keeps it bf16 NHWC
        for its Fcomb GEMMs; the reference's view of it is made when somebody reads the attribute."""
        v = self.__dict__.get("_unet_features")
        if callable(v):
            v = self.__dict__["_unet_features"] = v()
        return v

    @unet_features.setter
    def unet_features(self, v):
        self.__dict__["_unet_features"] = v

    def forward(self, patch, segm, training=True):
        """probabilistic_unet.py:215-223: sets posterior_latent_space (if training),
        prior_latent_space, unet_features; returns None."""
        patch = patch.contiguous().float()
        self._step = None
        if self._wants_grad():
            # ---- training path: train-mode BatchNorm, activations recorded for the backward kernels ----
            if not (self.training and training):
                raise NotImplementedError("gradients are built for net.train() + forward(training=True) (what train.py "
                                          "does); wrap inference in torch.no_grad()")
            if self.precision != "fp32" and (any(f % 64 for f in self.num_filters) or patch.shape[2] % 16 or patch.shape[3] % 16):
                raise NotImplementedError("tensor-core training needs channel counts that are multiples of 64 and H, W divisible "
                                          "by 16 (tcgen05 / TMA tiles); use precision='fp32' for this model")
            if segm is None:
                raise ValueError("forward(training=True) needs segm for the posterior")
            if patch.device.type != "cuda":
                raise RuntimeError("pmu_b200 models run on CUDA only (no CPU fallback)")
            st = train_engine.TrainStep(self, patch, segm.contiguous().float())
            self._step = st
            self._post, self._prior = (st.mu_q, st.ls_q), (st.mu_p, st.ls_p)
            self.posterior_latent_space = self._dist(st.mu_q, st.ls_q)
            self.prior_latent_space = self._dist(st.mu_p, st.ls_p)
            self.unet_features = st.features_nchw_f32 if st.tc_fcomb else st.feat     # cast on first read (property below)
            return
        self._enter("ProbabilisticUnet.forward")
        pk = self.packed()
        if training:
            if segm is None:
                raise ValueError("forward(training=True) needs segm for the posterior")
            mu_q, ls_q = pk.gaussian("posterior", patch, segm.contiguous().float())
            self._post = (mu_q, ls_q)
            self.posterior_latent_space = self._dist(mu_q, ls_q)
        mu_p, ls_p = pk.gaussian("prior", patch)
        self._prior = (mu_p, ls_p)
        self.prior_latent_space = self._dist(mu_p, ls_p)
        self.unet_features = pk.features_nchw_f32(pk.unet_features(patch))

    def sample(self, testing=False, *, z=None, eps=None):
        """probabilistic_unet.py:225-240: z ~ prior (rsample / sample) -> fcomb logits [B,C,H,W].
        Keyword-only z / eps inject the latent (z = mu + sigma*eps)."""
        step = getattr(self, "_step", None)
        if step is None:
            self._enter("ProbabilisticUnet.sample")
        if z is None:
            if eps is not None:
                z = self._prior[0] + torch.exp(self._prior[1]) * eps
            elif testing is False:
                z = self.prior_latent_space.rsample()
            else:
                z = self.prior_latent_space.sample()
        self.z_prior_sample = z
        if step is not None:
            # training step: the sample is only looked at (trainer.predict's return value never enters the loss,
            # probunet_trainer.py:27-39) — computed from the live weights, detached
            # (layer by layer through the register-tiled 1x1 kernels: 4x faster than the thread-per-pixel fused kernel)
            return train_engine.fcomb_forward(step, z.detach().float().contiguous())
        return self.packed().fcomb_logits(self.unet_features, z.float())

    def sample_at(self, z):
        """probabilistic_unet.py:242-247: fcomb at a given z ([L] -> broadcast over the batch)."""
        self._enter("ProbabilisticUnet.sample_at")
        z = z.to(self.unet_features.device, torch.float32)
        if z.dim() == 1:
            z = z.unsqueeze(0).expand(self.unet_features.shape[0], -1)
        return self.packed().fcomb_logits(self.unet_features, z.contiguous())

    def sample_grid(self, n_preds: int = 3, sigma_scale: float = 1.0, axes=(0, 1)):
        """Latent-grid sweep of visualize_sampling.py:21-31 in ONE fcomb launch: for every slice of the last forward(),
        z[i, j] = mu with z[axes[0]] = i * sigma_scale * sigma[axes[0]] + mu[axes[0]] and likewise axes[1] for j,
        i, j in range(-(n_preds // 2), n_preds // 2 + 1).  Returns (logits [B, G, G, C, H, W], z [B, G, G, L]);
        the reference loops G*G full `predict(slice, mask, z=z)` calls (each a whole U-Net pass) for the same result."""
        self._enter("ProbabilisticUnet.sample_grid")
        z = latent_grid_z(self._prior[0].float(), torch.exp(self._prior[1]).float(), n_preds, sigma_scale, axes)
        B, G = z.shape[0], z.shape[1]
        logits = self.packed().fcomb_logits(self.unet_features, z.reshape(B, G * G, -1))
        return logits.reshape(B, G, G, *logits.shape[2:]), z

    def reconstruct(self, use_posterior_mean=False, calculate_posterior=False, z_posterior=None):
        """probabilistic_unet.py:251-262."""
        self._enter("ProbabilisticUnet.reconstruct")
        if use_posterior_mean:
            z_posterior = self.posterior_latent_space.base_dist.loc
        elif calculate_posterior:
            z_posterior = self.posterior_latent_space.rsample()
        return self.packed().fcomb_logits(self.unet_features, z_posterior.float())

    def kl_divergence(self, analytic=True, calculate_posterior=False, z_posterior=None):
        """probabilistic_unet.py:264-279 -> [B]."""
        if analytic:
            return ops.kl_diag_gauss(self._post[0], self._post[1], self._prior[0], self._prior[1])
        if calculate_posterior:
            z_posterior = self.posterior_latent_space.rsample()
        return self.posterior_latent_space.log_prob(z_posterior) - self.prior_latent_space.log_prob(z_posterior)

    def elbo(self, segm, analytic_kl=True, reconstruct_posterior_mean=False, *, z=None, eps=None):
        """probabilistic_unet.py:281-308: -(sum CE + beta * mean_b KL); sets kl, reconstruction,
        reconstruction_loss."""
        step = getattr(self, "_step", None)
        if step is None:
            self._enter("ProbabilisticUnet.elbo")
        if self.n_classes == 1:
            raise NotImplementedError("num_classes == 1 ELBO is broken in the reference as well "
                                      "(probabilistic_unet.py:285-303, SURVEY.md App. B #8)")
        if z is not None:
            z_posterior = z
        elif eps is not None:
            z_posterior = self._post[0] + torch.exp(self._post[1]) * eps
        elif step is not None:
            # Normal.rsample() spelled out (loc + eps * scale with the same generator call), so eps is known to the backward
            eps = torch.distributions.utils._standard_normal(self._post[0].shape, dtype=torch.float32, device=self._post[0].device)
            z_posterior = self._post[0] + eps * torch.exp(self._post[1])
        else:
            z_posterior = self.posterior_latent_space.rsample()
        if step is not None:
            if reconstruct_posterior_mean:
                z_posterior, eps = self._post[0], torch.zeros_like(self._post[0])
            value = step.elbo(segm, z_posterior.float(), None if z is not None else eps, analytic_kl)
            self.kl, self.reconstruction, self.reconstruction_loss = step.kl, step.logits, step.rec
            return train_engine.elbo_with_grad(step, value, [p for p in self.parameters() if p.requires_grad])
        self.kl = torch.mean(self.kl_divergence(analytic=analytic_kl, calculate_posterior=False, z_posterior=z_posterior))
        self.reconstruction = self.reconstruct(use_posterior_mean=reconstruct_posterior_mean,
                                               calculate_posterior=False, z_posterior=z_posterior)
        self.reconstruction_loss = ops.ce_sum(self.reconstruction.contiguous(), segm.contiguous().float())
        return -(self.reconstruction_loss + self.beta * self.kl)
