#!/usr/bin/env python
"""Full-size fused outputs of BASELINE configs 2 and 3 from the CPU oracle, sampled on a voxel lattice.

    python -B tests/golden/make_golden_fullsize.py [2|3|both]

The oracle (oracle/pmu_oracle.py, pinned against the real reference by test_oracle_golden.py) runs the WHOLE
multi-planar prediction — every slice of every view, every latent sample — on the seeded synthetic inputs of
SURVEY.md §8d (phantom volume seed 1234, trainer model seed 0, eps = randn(3, D, N, 6) seed 4321), minutes of CPU
time, once.  Stored: mean / var [n,n,n,C] and entropy [n,n,n] on the lattice x, y, z in {o, o+step, ...} plus
whole-volume checksums, so the GPU test (tests/test_gpu_model.py::test_config{2,3}_full_size) can compare the FUSED
voxel-space results at the configurations' real sizes without the oracle having to run on the GPU box.

  config 2: 128^3, 8 samples   -> golden_cfg2_lattice.npz (lattice step 8, offset 3)
  config 3: 256^3, 16 samples  -> golden_cfg3_lattice.npz (lattice step 16, offset 5)
                               -> golden_cfg3_dense.npz: 1,048,576 voxels (x, y every 2nd from 1, z every 4th from 2):
                                  mean of classes 0 and 1 and entropy / ln 3 quantised to uint8 (half a step = 2e-3 on
                                  a probability), the fp32 argmax label and the oracle's top-2 margin (uint8) per voxel
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pmu_oracle as O  # noqa: E402

CFG = {"2": (128, 8, 8, 3), "3": (256, 16, 16, 5)}


def make(which):
    D, N, step, off = CFG[which]
    torch.set_num_threads(os.cpu_count())
    sd = O.make_state_dict(seed=0)
    vol, _ = O.phantom(D, seed=1234)
    eps = torch.randn(3, D, N, 6, generator=torch.Generator().manual_seed(4321))
    t = time.time()
    out = O.multiplanar_predict(vol, sd, eps, N, batch=8)
    dt = time.time() - t
    idx = torch.arange(off, D, step)
    lat = lambda v: v[idx][:, :, idx][:, :, :, idx]             # [x,C,y,z] -> [n,C,n,n]
    mean, var = lat(out["mean"]).permute(0, 2, 3, 1), lat(out["var"]).permute(0, 2, 3, 1)
    ent = out["entropy"][idx][:, idx][:, :, idx]
    path = os.path.join(HERE, f"golden_cfg{which}_lattice.npz")
    np.savez_compressed(path, D=D, N=N, step=step, offset=off, mean=mean.numpy(), var=var.numpy(), entropy=ent.numpy(),
                        mean_sum=out["mean"].double().sum((0, 2, 3)).numpy(), entropy_sum=float(out["entropy"].double().sum()),
                        var_sum=out["var"].double().sum((0, 2, 3)).numpy(),
                        labels_hist=np.bincount(out["mean"].argmax(1).flatten().numpy(), minlength=3))
    print(f"config {which}: {D}^3 x {N} samples, oracle {dt:.1f} s on {os.cpu_count()} threads -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KB)")
    if which == "3":
        ix, iz = torch.arange(1, D, 2), torch.arange(2, D, 4)
        m = out["mean"][ix][:, :, ix][:, :, :, iz]                  # [128, C, 128, 64]
        e = out["entropy"][ix][:, ix][:, :, iz]
        q8 = lambda v: torch.clamp(torch.round(v * 255.0), 0, 255).to(torch.uint8).numpy()
        top2 = torch.topk(m, 2, dim=1).values
        dpath = os.path.join(HERE, "golden_cfg3_dense.npz")
        np.savez_compressed(dpath, D=D, N=N, x0=1, xs=2, z0=2, zs=4, mean01_u8=q8(m[:, :2]), entropy_u8=q8(e / float(np.log(3.0))),
                            labels=m.argmax(1).to(torch.uint8).numpy(), margin_u8=q8(top2[:, 0] - top2[:, 1]))
        print(f"dense lattice: {m.shape[0] * m.shape[2] * m.shape[3]} voxels -> {dpath} ({os.path.getsize(dpath) / 1e6:.2f} MB)")


if __name__ == "__main__":
    w = sys.argv[1] if len(sys.argv) > 1 else "both"
    for k in (["2", "3"] if w == "both" else [w]):
        make(k)
