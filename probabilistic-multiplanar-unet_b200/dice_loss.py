"""dice_coeff drop-in (reference dice_loss.py:5-12): (2*sum(p*t)+1e-6)/(sum(p)+sum(t)+1e-6), the
sums taken over the WHOLE tensor.  The three sums run in one warp-reduced CUDA kernel
(pmu_dice_sums)."""
import torch

from . import ops


def dice_coeff(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    smooth = 0.000001
    s = ops.dice_sums(pred.contiguous().float(), target.contiguous().float())
    return (2.0 * s[0] + smooth) / (s[1] + s[2] + smooth)


def volume_dice(prob: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """eval.py:42-49 for every foreground class at once: prob [X,C,Y,Z] (avg_volume layout),
    truth [X,Y,Z] float labels -> dice [C-1] for k = 1..C-1."""
    smooth = 0.000001
    s = ops.argmax_dice_sums(prob.contiguous().float(), truth.contiguous().float())
    return (2.0 * s[:, 0] + smooth) / (s[:, 1] + s[:, 2] + smooth)
