#!/usr/bin/env python
"""eval.py — multi-planar evaluation entry point (drop-in for the reference's eval.py CLI).

    python eval.py -f CHECKPOINT -d DIR -m probunet [--samples 16] [--precision f16] [--out OUTDIR]

Same flags as the reference (eval.py:25-36): -f/--load a state_dict checkpoint, -d/--dir a folder
with images/ and labels/ (eval.py:96-98), -m/--model.  For every scan: predict labels along the
three standard views, compare each view's volume and the fused volume against the ground truth
(Dice of classes 1 and 2, eval.py:42-49,177-203), write the fused argmax label volume as NIfTI
(eval.py:51-57) and print mean / std per view (eval.py:218-233).

What differs from the (non-importable) reference script: the volume stays on the GPU, the network
runs ONCE per slice, N latent samples are drawn per slice and their PROBABILITIES are averaged
(the reference's 5-sample "average" never accumulates), BatchNorm is in eval
mode, and variance / entropy maps are written too.  Volumes are .nii / .nii.gz (built-in reader,
nibabel is not needed) or .npy.
"""
import argparse
import logging
import os

import numpy as np
import torch

import pmu_b200
from pmu_b200 import nifti_io


def get_args():
    parser = argparse.ArgumentParser(description="Predict using a trained ProbabilisticUnet",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("-f", "--load", dest="load", type=str, default=False, help="Load model from a .pth file")
    parser.add_argument("-d", "--dir", dest="dir", type=str, default=None, help="image and label superdirs.")
    parser.add_argument("-m", "--model", dest="net", type=str, default="probunet", help="what model to use: unet or probunet")
    parser.add_argument("--samples", type=int, default=16, help="latent samples per slice")
    parser.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32"])
    parser.add_argument("--slice-batch", type=int, default=64)
    parser.add_argument("--seed", type=int, default=4321)
    parser.add_argument("--out", type=str, default="predictions", help="where the label / uncertainty volumes go")
    return parser.parse_args()


def load_volume(path: str) -> np.ndarray:
    if path.endswith(".npy"):
        return np.load(path).astype(np.float64)
    return nifti_io.load(path)


def main():
    logging.basicConfig(level=logging.INFO, format="%(levelname)s: %(message)s")
    args = get_args()
    if args.net != "probunet":
        raise SystemExit("Error! only the probabilistic model ('-m probunet') is on the B200 path")
    if not torch.cuda.is_available():
        raise SystemExit("eval.py needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda")
    logging.info(f"Using device {device}")
    train = pmu_b200.ProbUNetTrainer(device, n_channels=1, n_classes=3, load_model=args.load or None, latent_dim=6,
                                     beta=10, precision=args.precision)
    train.net.eval()
    dir_img, dir_mask = os.path.join(args.dir, "images"), os.path.join(args.dir, "labels")
    ids = sorted(os.listdir(dir_img))
    os.makedirs(args.out, exist_ok=True)
    pred = pmu_b200.MultiPlanarPredictor(train.net, device, precision=args.precision, n_samples=args.samples,
                                         slice_batch=args.slice_batch)
    view_dice, fused_dice = [[], [], []], []
    for idx in ids:
        vol = load_volume(os.path.join(dir_img, idx)).astype(np.float32)
        out = pred.predict(vol, seed=args.seed, want_labels=True, per_plane=True)
        stem = idx.split(".")[0]
        nifti_io.save(os.path.join(args.out, stem + "_labels.nii"), out["labels"].cpu().numpy())      # eval.py:51-57
        nifti_io.save(os.path.join(args.out, stem + "_entropy.nii"), out["entropy"].cpu().numpy())
        mpath = os.path.join(dir_mask, idx)
        if os.path.exists(mpath):
            truth = load_volume(mpath).astype(np.float32)
            pd = pmu_b200.padded_dims(truth.shape)
            if pd != truth.shape:
                big = np.zeros(pd, np.float32); big[:truth.shape[0], :truth.shape[1], :truth.shape[2]] = truth; truth = big
            t = torch.from_numpy(truth).to(device)
            for v in range(3):
                view_dice[v].append(pmu_b200.volume_dice(out["plane_means"][v], t).cpu().numpy())
            fused_dice.append(pmu_b200.volume_dice(out["mean"], t).cpu().numpy())
            logging.info(f"{idx}: fused dice {fused_dice[-1]}")
    if fused_dice:
        for v in range(3):
            a = np.array(view_dice[v])
            print(f"view {v + 1} dice: mean={a.mean(0)}, std={a.std(0)}")
        a = np.array(fused_dice)
        print(f"avg volume: mean={a.mean(0)}, std={a.std(0)}")


if __name__ == "__main__":
    main()
