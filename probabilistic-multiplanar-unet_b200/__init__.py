"""pmu_b200 — B200-native multi-planar probabilistic U-Net inference path.

Import as ``pmu_b200`` (the directory name carries the reference's hyphenated name; the
top-level ``pmu_b200`` package is an alias whose __path__ points here).
"""
from . import _lib, nifti_io, ops  # noqa: F401
from .dice_loss import dice_coeff, volume_dice  # noqa: F401
from .engine import PackedNet  # noqa: F401
from .model import ProbabilisticUnet, UNet  # noqa: F401
from .mri_dataset import MRI_Dataset, view_affine  # noqa: F401
from .multiplanar import (MultiPlanarPredictor, padded_dims, reduce_accumulators, reduce_scatter_accumulators,  # noqa: F401
                          shard_slices)
from .trainer import ProbUNetTrainer  # noqa: F401
from .train_dp import allreduce_gradients, dp_train_step, sync_batchnorm_buffers  # noqa: F401

__version__ = "0.1.0"
