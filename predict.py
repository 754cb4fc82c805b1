#!/usr/bin/env python
"""predict.py — the reference's predict() hook (predict.py:15-19) and a one-volume CLI.

    python predict.py -f CHECKPOINT -i VOLUME.nii[.gz] [-o OUTDIR] [--samples 16] [--precision f16]
"""
import argparse
import os

import numpy as np
import torch

import pmu_b200
from pmu_b200 import nifti_io


def predict(net, imgs, masks, train=True, prob=False):
    """Reference signature (predict.py:15-19): forward + one prior sample; returns the logits (the
    reference stub forgets the return)."""
    if prob:
        with torch.set_grad_enabled(False):
            net.forward(imgs, masks, training=train)
            return net.sample(testing=(not train))
    raise NotImplementedError("only the probabilistic model is on the B200 path (prob=True)")


def main():
    ap = argparse.ArgumentParser(description="Multi-planar probabilistic prediction of one volume")
    ap.add_argument("-f", "--load", type=str, default=None)
    ap.add_argument("-i", "--input", type=str, required=True)
    ap.add_argument("-o", "--out", type=str, default="predictions")
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32"])
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("predict.py needs a CUDA device (there is no CPU fallback)")
    tr = pmu_b200.ProbUNetTrainer(torch.device("cuda"), 1, 3, load_model=args.load, latent_dim=6, beta=10,
                                  precision=args.precision)
    tr.net.eval()
    vol = (np.load(args.input) if args.input.endswith(".npy") else nifti_io.load(args.input)).astype(np.float32)
    out = pmu_b200.MultiPlanarPredictor(tr.net, "cuda", precision=args.precision, n_samples=args.samples).predict(
        vol, want_labels=True)
    os.makedirs(args.out, exist_ok=True)
    stem = os.path.basename(args.input).split(".")[0]
    nifti_io.save(os.path.join(args.out, stem + "_labels.nii"), out["labels"].cpu().numpy())
    nifti_io.save(os.path.join(args.out, stem + "_entropy.nii"), out["entropy"].cpu().numpy())
    nifti_io.save(os.path.join(args.out, stem + "_variance.nii"), out["var"].sum(1).cpu().numpy())
    print(f"wrote {args.out}/{stem}_{{labels,entropy,variance}}.nii")


if __name__ == "__main__":
    main()
