"""CPU oracle for the multi-planar probabilistic inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``pmu_b200``) never imports this module.

This is a *restatement* (plain torch-CPU / numpy, fp32) of the reference
``qzs634/Probabilistic-Multiplanar-Unet`` algorithm for the hot path.  Every
function cites the reference file:line it follows (paths relative to
``Probabilistic-Multiplanar-Unet/`` in the reference tree).

Parity pin — PINNED against the reference's own code run in the build container.  The reference has no
tests and no golden vectors of its own, so the vectors are generated from the REAL
reference and committed under ``tests/golden/`` together with the scripts that made them;
``tests/test_oracle_golden.py`` checks this restatement against every one of them:
  * model half — ``make_golden.py`` / ``make_golden_small_extra.py`` / ``make_golden_trainer_api.py`` /
    ``make_golden_grads.py`` import the real ``model/``, ``dice_loss.py`` and ``trainer/`` modules (two import shims,
    SURVEY App. C): forward / sample_at / reconstruct / analytic and Monte-Carlo KL / ELBO in eval- and train-mode
    BatchNorm, the trainer architecture, ``ProbUNetTrainer.predict / eval / loss / mask_to_image``, ``dice_coeff``,
    and every parameter gradient of ``loss.backward()``;
  * data-plane half — ``utils/mri_dataset.py`` is executed unmodified with stub ``nibabel`` / ``utils.dataset``
    modules (its only unmet imports) by ``make_golden_dataplane.py``: index map, ``pad_dimensions``,
    ``sample_slice``, ``preprocess``, ``__getitem__`` match bit for bit; ``eval.py`` cannot be parsed as a whole
    (syntax error at :137-138), so its ``dice``, ``slices_to_volume``, reassembly / fusion statements and the
    per-slice sample loop are exec'd verbatim by line range (``make_golden_dataplane.py``,
    ``make_golden_sampleloop.py``); the latent-grid loop of ``visualize_sampling.py:21-31`` likewise
    (``make_golden_latent_grid.py``).
Build-defined and therefore WITHOUT a reference counterpart (SURVEY.md Appendix A is their spec): N-sample
averaging of probabilities, per-voxel variance and entropy, nearest / trilinear resampling onto arbitrary grids.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------
# Data plane in: views, padding, slicing, normalisation  (utils/mri_dataset.py)
# ----------------------------------------------------------------------------


def standard_views() -> List[np.ndarray]:
    """The three axis-aligned view vectors (mri_dataset.py:60-66)."""
    return [np.array([1, 0, 0]), np.array([0, 1, 0]), np.array([0, 0, 1])]


def pad_dimensions(vol: np.ndarray) -> np.ndarray:
    """Zero-pad the arg-min axis at its high end by (max-min) (mri_dataset.py:85-98).

    Only ONE axis is padded, exactly like the reference (a volume with two short
    axes stays non-cubic).
    """
    shape = vol.shape
    diff = int(np.max(shape) - np.min(shape))
    if diff == 0:
        return vol
    axis = int(np.argmin(shape))
    pad = [(0, 0)] * 3
    pad[axis] = (0, diff)
    return np.pad(vol, pad, mode="constant", constant_values=0)


def index_map(dims: Sequence[int]) -> List[Tuple[int, int]]:
    """Flat (view, slice) order for ONE scan with filter=False (mri_dataset.py:37-49)."""
    out = []
    for view in range(3):
        for s in range(int(dims[view])):
            out.append((view, s))
    return out


def sample_slice(vol: np.ndarray, view: int, s: int) -> np.ndarray:
    """vol[s,:,:] / vol[:,s,:] / vol[:,:,s] (mri_dataset.py:70-82)."""
    if view == 0:
        return vol[s, :, :]
    if view == 1:
        return vol[:, s, :]
    if view == 2:
        return vol[:, :, s]
    raise ValueError("No valid view")


def preprocess(img: np.ndarray, label: bool = False) -> np.ndarray:
    """Add channel dim (HWC->CHW) and divide the image slice by its own max if
    max != 0 (mri_dataset.py:101-112).  The divide happens in the array's dtype;
    the reference holds fp64 (nibabel get_fdata) and casts to fp32 afterwards
    (mri_dataset.py:142), which is what ``normalised_slice`` reproduces."""
    if img.ndim == 2:
        img = img[:, :, None]
    t = np.transpose(img, (2, 0, 1))
    if not label:
        m = np.max(t)
        if not m == 0:
            t = t / m
    return t


def normalised_slice(vol: np.ndarray, view: int, s: int) -> np.ndarray:
    """[1,H,W] fp32 slice exactly as MRI_Dataset.__getitem__ hands it to the net
    (mri_dataset.py:134-142): fp64 divide by the slice max, then cast to fp32."""
    sl = sample_slice(vol, view, s).astype(np.float64)
    return preprocess(sl, label=False).astype(np.float32)


def plane_slices(vol: np.ndarray, view: int, s0: int = 0, ns: Optional[int] = None,
                 normalise: bool = True) -> np.ndarray:
    """Batch of slices [ns,1,H,W] fp32 for one plane."""
    if ns is None:
        ns = vol.shape[view] - s0
    outs = []
    for s in range(s0, s0 + ns):
        if normalise:
            outs.append(normalised_slice(vol, view, s))
        else:
            outs.append(sample_slice(vol, view, s).astype(np.float32)[None])
    return np.stack(outs, 0)


# ----------------------------------------------------------------------------
# [build-defined] resampling onto an affine slice grid
# ----------------------------------------------------------------------------

def identity_affine(view: int) -> np.ndarray:
    """12 floats [o(3), n(3), u(3), v(3)]: q = o + s*n + r*u + c*v reproduces
    sample_slice(view) exactly (mri_dataset.py:72-77)."""
    e = np.eye(3, dtype=np.float32)
    o = np.zeros(3, np.float32)
    if view == 0:
        n, u, v = e[0], e[1], e[2]
    elif view == 1:
        n, u, v = e[1], e[0], e[2]
    else:
        n, u, v = e[2], e[0], e[1]
    return np.concatenate([o, n, u, v]).astype(np.float32)


def _c_oracle():
    """The C part of the oracle (oracle/resample_fma.c: the resampling spec needs correctly rounded fused multiply-adds,
    which numpy cannot express), compiled with gcc into oracle/_build/ on first use and loaded with ctypes."""
    global _C_LIB
    if _C_LIB is not None:
        return _C_LIB
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "resample_fma.c")
    out_dir = os.path.join(here, "_build")
    lib = os.path.join(out_dir, "libpmu_oracle.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        tmp = lib + f".{os.getpid()}.tmp"
        # -ffp-contract=off: the ONLY fused operations are the explicit fmaf() calls of the spec
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", src, "-o", tmp, "-lm"], check=True)
        os.replace(tmp, lib)
    L = ctypes.CDLL(lib)
    fp = ctypes.POINTER(ctypes.c_float)
    L.pmu_oracle_resample.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, fp]
    L.pmu_oracle_resample.restype = None
    L.pmu_oracle_scatter_nearest.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp]
    L.pmu_oracle_scatter_nearest.restype = None
    _C_LIB = L
    return L


_C_LIB = None


def _fptr(a: np.ndarray):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def resample_slices(vol: np.ndarray, affine: np.ndarray, s0: int, ns: int, H: int, W: int,
                    mode: str = "nearest") -> np.ndarray:
    """[ns,H,W] fp32 raw (un-normalised) resampled slices, zeros outside the volume.

    Spec (oracle/resample_fma.c, every operation a correctly rounded fp32 operation):
      q[ax] = fma(c, v[ax], fma(r, u[ax], fma(s, n[ax], o[ax])))
      nearest  : index = floor(q + 0.5) per axis.
      trilinear: 8 taps, lerp along z, then y, then x, lerp(a, b, t) = fma(t, b - a, a) (align_corners=True convention,
                 voxel centres at integer coordinates).
    On the identity grid both modes reproduce plain slicing bit-exactly."""
    if mode not in ("nearest", "trilinear"):
        raise ValueError(mode)
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    aff = np.ascontiguousarray(np.asarray(affine, dtype=np.float32).reshape(12))
    out = np.empty((ns, H, W), dtype=np.float32)
    _c_oracle().pmu_oracle_resample(_fptr(vol), vol.shape[0], vol.shape[1], vol.shape[2], _fptr(aff), int(s0), int(ns), int(H),
                                    int(W), int(mode == "trilinear"), _fptr(out))
    return out


def scatter_nearest(vals: np.ndarray, affine: np.ndarray, s0: int, dims: Sequence[int], acc: np.ndarray, cnt: np.ndarray) -> None:
    """[build-defined] App. A step 6 for a non-identity grid: vals [ns,K,H,W] are added to acc [d0,K,d1,d2] at the voxel
    NEAREST to every output pixel (floor(q + 0.5)), cnt [d0,d1,d2] counts the contributions; pixels outside the volume
    are dropped.  In place."""
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    aff = np.ascontiguousarray(np.asarray(affine, dtype=np.float32).reshape(12))
    assert acc.dtype == np.float32 and cnt.dtype == np.float32 and acc.flags.c_contiguous and cnt.flags.c_contiguous
    ns, K, H, W = vals.shape
    _c_oracle().pmu_oracle_scatter_nearest(_fptr(vals), K, _fptr(aff), int(s0), ns, H, W, int(dims[0]), int(dims[1]), int(dims[2]),
                                           _fptr(acc), _fptr(cnt))


def normalise_slices(raw: np.ndarray) -> np.ndarray:
    """Per-slice x/max(x) if max != 0, fp64 divide then fp32 cast (mri_dataset.py:108-110,142)."""
    out = np.empty_like(raw, dtype=np.float32)
    for i in range(raw.shape[0]):
        m = np.max(raw[i])
        if m != 0:
            out[i] = (raw[i].astype(np.float64) / np.float64(m)).astype(np.float32)
        else:
            out[i] = raw[i]
    return out


# ----------------------------------------------------------------------------
# Model: functional restatement over a reference-schema state_dict
# ----------------------------------------------------------------------------

def _bn(x, sd: StateDict, p: str, train: bool, eps: float = 1e-5):
    """nn.BatchNorm2d (unet_parts.py:16,19; probabilistic_unet.py:39,44).  ``train``
    uses batch statistics (biased variance) like a module in train() mode; running
    stats are NOT updated here (the oracle is stateless)."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    if train:
        return F.batch_norm(x, None, None, w, b, True, 0.0, eps)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], w, b, False, 0.0, eps)


def _conv_bn_relu(x, sd: StateDict, conv: str, bn: str, train: bool):
    x = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=1)
    return F.relu(_bn(x, sd, bn, train))


def _double_conv(x, sd: StateDict, p: str, train: bool):
    """DoubleConv = (conv3x3 pad1 -> BN -> ReLU) x2 (unet_parts.py:9-24)."""
    x = _conv_bn_relu(x, sd, p + ".double_conv.0", p + ".double_conv.1", train)
    return _conv_bn_relu(x, sd, p + ".double_conv.3", p + ".double_conv.4", train)


def num_levels(sd: StateDict, prefix: str = "unet") -> int:
    n = 0
    while f"{prefix}.down_blocks.{n}.maxpool_conv.1.double_conv.0.weight" in sd:
        n += 1
    return n + 1


def unet_features(sd: StateDict, x: torch.Tensor, bn_train: bool = False, prefix: str = "unet") -> torch.Tensor:
    """UNet.forward with apply_last_layer=False: returns the last decoder map
    (unet_model.py:31-54).  inc -> Down x(L-1) (MaxPool2d(2)+DoubleConv,
    unet_parts.py:27-38) -> Up x(L-1) (ConvTranspose2d k2 s2, F.pad, cat([skip, up]),
    DoubleConv; unet_parts.py:41-67).  up_blocks are stored reversed
    (unet_model.py:29): up_blocks.0 is the deepest."""
    L = num_levels(sd, prefix)
    xs = [_double_conv(x, sd, f"{prefix}.inc", bn_train)]
    for i in range(L - 1):
        h = F.max_pool2d(xs[i], 2)
        xs.append(_double_conv(h, sd, f"{prefix}.down_blocks.{i}.maxpool_conv.1", bn_train))
    h = xs[-1]
    for i in range(L - 1):
        skip = xs[L - 2 - i]
        up = F.conv_transpose2d(h, sd[f"{prefix}.up_blocks.{i}.up.weight"], sd[f"{prefix}.up_blocks.{i}.up.bias"], stride=2)
        dy, dx = skip.shape[2] - up.shape[2], skip.shape[3] - up.shape[3]
        up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        h = _double_conv(torch.cat([skip, up], 1), sd, f"{prefix}.up_blocks.{i}.conv", bn_train)
    return h


def unet_outc(sd: StateDict, feat: torch.Tensor, prefix: str = "unet") -> torch.Tensor:
    """OutConv 1x1 (unet_parts.py:70-76) + sigmoid when n_classes==1 (unet_model.py:48-49)."""
    w = sd[f"{prefix}.outc.conv.weight"]
    out = F.conv2d(feat, w, sd[f"{prefix}.outc.conv.bias"])
    return torch.sigmoid(out) if w.shape[0] == 1 else out


def encoder(sd: StateDict, prefix: str, x: torch.Tensor, bn_train: bool = False) -> torch.Tensor:
    """Encoder (probabilistic_unet.py:11-53): level i>0 starts with
    AvgPool2d(2,2,ceil_mode=True); each level = 2x (conv3x3 pad1 + BN + ReLU).
    Sequential indices: conv1 = 7i, bn1 = 7i+1, conv2 = 7i+3, bn2 = 7i+4."""
    i = 0
    while f"{prefix}.layers.{7 * i}.weight" in sd:
        if i > 0:
            x = F.avg_pool2d(x, 2, 2, 0, ceil_mode=True)
        x = _conv_bn_relu(x, sd, f"{prefix}.layers.{7 * i}", f"{prefix}.layers.{7 * i + 1}", bn_train)
        x = _conv_bn_relu(x, sd, f"{prefix}.layers.{7 * i + 3}", f"{prefix}.layers.{7 * i + 4}", bn_train)
        i += 1
    return x


def gaussian_head(sd: StateDict, which: str, x: torch.Tensor, segm: Optional[torch.Tensor] = None,
                  bn_train: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """AxisAlignedConvGaussian.forward (probabilistic_unet.py:82-114): optional
    cat(input, segm) (posterior), encoder, mean over H then W, 1x1 conv to 2L,
    mu = [:, :L], log_sigma = [:, L:].  Returns (mu, log_sigma); sigma = exp(log_sigma)."""
    if segm is not None:
        x = torch.cat((x, segm), 1)
    e = encoder(sd, f"{which}.encoder", x, bn_train)
    e = torch.mean(e, dim=2, keepdim=True)
    e = torch.mean(e, dim=3, keepdim=True)
    ml = F.conv2d(e, sd[f"{which}.conv_layer.weight"], sd[f"{which}.conv_layer.bias"])[:, :, 0, 0]
    L = ml.shape[1] // 2
    return ml[:, :L], ml[:, L:]


def fcomb(sd: StateDict, feat: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """Fcomb.forward (probabilistic_unet.py:155-181): broadcast z[B,L] over HxW,
    cat((features, z), 1), [1x1 conv + ReLU] x (no_convs_fcomb-1), last 1x1 conv -> logits."""
    B, _, H, W = feat.shape
    zt = z[:, :, None, None].expand(B, z.shape[1], H, W)
    h = torch.cat((feat, zt), 1)
    i = 0
    while f"fcomb.layers.{2 * i}.weight" in sd:
        h = F.relu(F.conv2d(h, sd[f"fcomb.layers.{2 * i}.weight"], sd[f"fcomb.layers.{2 * i}.bias"]))
        i += 1
    return F.conv2d(h, sd["fcomb.last_layer.weight"], sd["fcomb.last_layer.bias"])


def kl_diag_gauss(mu_q, log_sigma_q, mu_p, log_sigma_p) -> torch.Tensor:
    """Analytic KL(q||p) of diagonal Gaussians, summed over the latent dim -> [B]
    (probabilistic_unet.py:264-279 via torch.distributions.kl for Independent(Normal)):
    sum_l [ log(sp/sq) + (sq^2 + (mq-mp)^2)/(2 sp^2) - 1/2 ]."""
    sq, sp = torch.exp(log_sigma_q), torch.exp(log_sigma_p)
    var_ratio = (sq / sp) ** 2
    t1 = ((mu_q - mu_p) / sp) ** 2
    return (0.5 * (var_ratio + t1 - 1 - torch.log(var_ratio))).sum(1)


def kl_monte_carlo(mu_q, log_sigma_q, mu_p, log_sigma_p, z) -> torch.Tensor:
    """One-sample estimate log q(z) - log p(z) -> [B] (probabilistic_unet.py:274-278, kl_divergence(analytic=False)):
    log N(z; mu, sigma) summed over the latent dim = sum_l [ -((z-mu)/sigma)^2 / 2 - log sigma - log sqrt(2 pi) ]
    (Independent(Normal(mu, exp(log_sigma)), 1).log_prob); the 2 pi terms cancel in the difference."""
    def log_prob(mu, ls):
        return (-0.5 * ((z - mu) / torch.exp(ls)) ** 2 - ls).sum(1)
    return log_prob(mu_q, log_sigma_q) - log_prob(mu_p, log_sigma_p)


def ce_sum(logits: torch.Tensor, segm: torch.Tensor) -> torch.Tensor:
    """CrossEntropyLoss(reduction none) summed over batch and pixels
    (probabilistic_unet.py:288,300-304).  segm: float [B,1,H,W] integer labels."""
    tgt = segm.to(torch.long).squeeze(1)
    return F.cross_entropy(logits, tgt, reduction="none").sum()


def elbo(sd: StateDict, patch, segm, eps_q, beta: float, bn_train: bool = True, z_q=None):
    """ProbabilisticUnet.forward(training=True) + elbo(segm) with the posterior
    reparameterisation noise injected: z_q = mu_q + sigma_q * eps_q
    (probabilistic_unet.py:215-223, 281-308).  Returns dict with kl (mean over
    batch), reconstruction_loss (CE sum), elbo = -(rec + beta*kl), logits."""
    mu_q, ls_q = gaussian_head(sd, "posterior", patch, segm, bn_train)
    mu_p, ls_p = gaussian_head(sd, "prior", patch, None, bn_train)
    feat = unet_features(sd, patch, bn_train)
    if z_q is None:
        z_q = mu_q + torch.exp(ls_q) * eps_q
    kl = kl_diag_gauss(mu_q, ls_q, mu_p, ls_p).mean()
    logits = fcomb(sd, feat, z_q)
    rec = ce_sum(logits, segm)
    return {"kl": kl, "reconstruction_loss": rec, "elbo": -(rec + beta * kl), "logits": logits,
            "mu_q": mu_q, "log_sigma_q": ls_q, "mu_p": mu_p, "log_sigma_p": ls_p, "features": feat}


def dice_coeff(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """(2*sum(p*t)+1e-6)/(sum(p)+sum(t)+1e-6) over the WHOLE tensor (dice_loss.py:5-12)."""
    smooth = 0.000001
    inter = (pred.reshape(-1) * target.reshape(-1)).sum()
    return (2.0 * inter + smooth) / (pred.sum() + target.sum() + smooth)


def argmax_dice(prob: torch.Tensor, truth: torch.Tensor, k: int) -> float:
    """eval.py:42-49 / probunet_trainer.py:41-60: dice of one-hot(argmax over dim 1)[:,k]
    against (truth == k)."""
    idx = torch.argmax(prob, 1, keepdim=True)
    one_hot = torch.zeros_like(prob).scatter_(1, idx, 1.0)
    return float(dice_coeff(one_hot[:, k], (truth == k).float().reshape(one_hot[:, k].shape)))


# ----------------------------------------------------------------------------
# Data plane out: softmax, reassembly, fusion  (eval.py:157,162-193 + App. A 5-7)
# ----------------------------------------------------------------------------

def eval_sample_loop(draws: Sequence[torch.Tensor], literal: bool = True) -> torch.Tensor:
    """The per-slice sample loop + softmax of eval.py:145-157 on the logit maps `draws` that five stochastic
    ProbUNetTrainer.predict calls returned.  literal=True is what the file computes: the `+` on eval.py:152 discards its
    result, so pred_masks = draws[0] / len(draws) and probs = softmax(draws[0] / 5, dim=1).
    literal=False is the evident intent — softmax of the MEAN logits.  (The build's N-sample fusion averages
    probabilities, App. A step 5; this function exists to pin row a20 of SURVEY.md §8 against the reference.)"""
    n = len(draws)
    pred = draws[0].clone() if literal else torch.stack(list(draws)).sum(0)
    pred = pred / n
    return torch.softmax(pred, dim=1)


def scatter_plane(view: int, per_slice: torch.Tensor) -> torch.Tensor:
    """[S,C,H,W] per-slice maps of one view -> [x,C,y,z] volume.  view 0: as is
    (eval.py:176); view 1: permute(2,1,0,3) (eval.py:182); view 2: permute(2,1,3,0)
    (eval.py:188)."""
    if view == 0:
        return per_slice
    if view == 1:
        return per_slice.permute(2, 1, 0, 3)
    return per_slice.permute(2, 1, 3, 0)


def finalize(S1: torch.Tensor, S2: torch.Tensor, count: float):
    """[build-defined, App. A step 7] mean = S1/count; var = max(S2/count - mean^2, 0);
    entropy = sum_k -mean_k ln mean_k (0 at mean_k = 0).  Layout [x,C,y,z]; entropy [x,y,z].
    For count = 3, N = 1 the mean equals eval.py:193's (v1+v2+v3)/3."""
    inv = 1.0 / float(count)
    mean = S1 * inv
    var = torch.clamp(S2 * inv - mean * mean, min=0.0)
    ent = torch.special.entr(mean).sum(1)
    return mean, var, ent


@torch.no_grad()
def multiplanar_predict(vol: np.ndarray, sd: StateDict, eps: torch.Tensor, n_samples: int,
                        planes: Sequence[int] = (0, 1, 2), batch: int = 8,
                        slice_ranges: Optional[Dict[int, Tuple[int, int]]] = None,
                        return_per_slice: bool = False):
    """App. A steps 1-7 on a cubic fp32 volume with eval-mode BN and injected
    latents: eps[P, D, N, L];  z[p,s,n] = mu[p,s] + sigma[p,s] * eps[p,s,n].

    Returns dict(S1, S2 [x,C,y,z] fp32 sums over all planes/samples, mean, var,
    entropy, count).  ``slice_ranges`` restricts plane p to slices [s0, s1) (used
    to check multi-GPU sharding: partial sums add up)."""
    vol = pad_dimensions(np.asarray(vol))
    D = vol.shape
    C = sd["fcomb.last_layer.weight"].shape[0]
    S1 = torch.zeros(D[0], C, D[1], D[2])
    S2 = torch.zeros_like(S1)
    per_slice = {}
    for p in planes:
        s_lo, s_hi = (0, D[p]) if slice_ranges is None else slice_ranges.get(p, (0, 0))
        if s_hi <= s_lo:
            continue
        H, W = [D[a] for a in range(3) if a != p]
        p1 = torch.zeros(s_hi - s_lo, C, H, W)
        p2 = torch.zeros_like(p1)
        for b0 in range(s_lo, s_hi, batch):
            nb = min(batch, s_hi - b0)
            x = torch.from_numpy(plane_slices(vol, p, b0, nb))
            feat = unet_features(sd, x)
            mu, ls = gaussian_head(sd, "prior", x)
            sigma = torch.exp(ls)
            for n in range(n_samples):
                z = mu + sigma * eps[p, b0:b0 + nb, n]
                prob = torch.softmax(fcomb(sd, feat, z), dim=1)
                p1[b0 - s_lo:b0 - s_lo + nb] += prob
                p2[b0 - s_lo:b0 - s_lo + nb] += prob * prob
        full1 = torch.zeros(D[p], C, H, W)
        full2 = torch.zeros_like(full1)
        full1[s_lo:s_hi] = p1
        full2[s_lo:s_hi] = p2
        S1 += scatter_plane(p, full1)
        S2 += scatter_plane(p, full2)
        if return_per_slice:
            per_slice[p] = (p1, p2)
    count = float(len(planes) * n_samples)
    mean, var, ent = finalize(S1, S2, count)
    out = {"S1": S1, "S2": S2, "mean": mean, "var": var, "entropy": ent, "count": count}
    if return_per_slice:
        out["per_slice"] = per_slice
    return out


# ----------------------------------------------------------------------------
# Seeded synthetic inputs shared by tests and bench
# ----------------------------------------------------------------------------

def phantom(D: int, seed: int = 1234, dims: Optional[Sequence[int]] = None):
    """vol = 0.6*phantom + 0.4*rand, phantom = two nested offset ellipsoids
    (1.0 outer shell / 0.5 inner... see labels); labels 0 background, 1 shell, 2 core."""
    dims = tuple(dims) if dims is not None else (D, D, D)
    g = torch.Generator().manual_seed(seed)
    ax = [torch.linspace(-1, 1, n) for n in dims]
    X, Y, Z = torch.meshgrid(*ax, indexing="ij")
    outer = ((X - 0.05) / 0.80) ** 2 + ((Y + 0.10) / 0.65) ** 2 + (Z / 0.70) ** 2 <= 1.0
    inner = ((X - 0.15) / 0.40) ** 2 + ((Y + 0.05) / 0.30) ** 2 + ((Z - 0.10) / 0.35) ** 2 <= 1.0
    ph = torch.zeros(dims)
    ph[outer] = 1.0
    ph[inner] = 0.5
    labels = torch.zeros(dims)
    labels[outer] = 1.0
    labels[inner] = 2.0
    vol = (0.6 * ph + 0.4 * torch.rand(dims, generator=g)).float()
    return vol.numpy(), labels.numpy()


def make_state_dict(num_filters: Sequence[int] = (64, 128, 256, 512, 1024), input_channels: int = 1,
                    num_classes: int = 3, latent_dim: int = 6, no_convs_fcomb: int = 4,
                    seed: int = 0) -> StateDict:
    """Reference-schema state_dict (key names/shapes of ProbabilisticUnet.state_dict(),
    probabilistic_unet.py:194-213, unet_model.py:10-29) filled with seeded values:
    kaiming-normal conv weights, small biases, and NON-trivial BN affine + running
    statistics so that BN folding is exercised (SURVEY.md §8d "Weights")."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}

    def conv(name, cout, cin, k, transposed=False):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        fan_in = shape[1] * k * k
        sd[name + ".weight"] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.01

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    def dconv(p, cin, cout):
        conv(p + ".double_conv.0", cout, cin, 3); bn(p + ".double_conv.1", cout)
        conv(p + ".double_conv.3", cout, cout, 3); bn(p + ".double_conv.4", cout)

    nf = list(num_filters)
    L = len(nf)
    for i in range(L - 1):
        dconv(f"unet.down_blocks.{i}.maxpool_conv.1", nf[i], nf[i + 1])
    for i in range(L - 1):
        cin = nf[L - 1 - i]
        conv(f"unet.up_blocks.{i}.up", cin // 2, cin, 2, transposed=True)
        dconv(f"unet.up_blocks.{i}.conv", cin, cin // 2)
    dconv("unet.inc", input_channels, nf[0])
    conv("unet.outc.conv", num_classes, nf[0], 1)
    for which, extra in (("prior", 0), ("posterior", 1)):
        for i in range(L):
            cin = input_channels + extra if i == 0 else nf[i - 1]
            conv(f"{which}.encoder.layers.{7 * i}", nf[i], cin, 3); bn(f"{which}.encoder.layers.{7 * i + 1}", nf[i])
            conv(f"{which}.encoder.layers.{7 * i + 3}", nf[i], nf[i], 3); bn(f"{which}.encoder.layers.{7 * i + 4}", nf[i])
        conv(f"{which}.conv_layer", 2 * latent_dim, nf[-1], 1)
        # keep log_sigma moderate so exp() stays well-conditioned
        sd[f"{which}.conv_layer.bias"] = torch.randn(2 * latent_dim, generator=g) * 0.3
    conv("fcomb.layers.0", nf[0], nf[0] + latent_dim, 1)
    for j in range(1, no_convs_fcomb - 1):
        conv(f"fcomb.layers.{2 * j}", nf[0], nf[0], 1)
    conv("fcomb.last_layer", num_classes, nf[0], 1)
    return sd
