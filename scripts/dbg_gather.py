import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import numpy as np, torch
from oracle import pmu_oracle as O
from pmu_b200 import ops
affs = [O.identity_affine(p) for p in range(3)]
affs.append(np.array([0.3, -0.2, 0.4, 0.9, 0.1, 0.0, -0.1, 0.95, 0.05, 0.02, 0.0, 1.05], np.float32))
affs.append(np.array([-2.5, 3.0, 1.0, 1.5, 0.0, 0.0, 0.0, 0.5, 0.25, 0.0, -0.25, 0.5], np.float32))
vol2, _ = O.phantom(0, seed=4, dims=(48, 40, 64))
v2 = torch.from_numpy(vol2).cuda()
for mode in ("nearest", "trilinear"):
    for i, aff in enumerate(affs):
        H, W = [(40, 64), (48, 64), (48, 40), (40, 64), (37, 61)][i]
        for (s0, ns) in [(0, 40), (3, 33)]:
            for wm in (False, True):
                try:
                    r = ops.slice_gather(v2, 0, s0, ns, interp=mode, affine=aff, hw=(H, W), want_max=wm)
                    torch.cuda.synchronize()
                    got = (r[0] if wm else r).cpu().numpy()[:, 0]
                    ref = O.resample_slices(vol2, aff, s0, ns, H, W, mode)
                    print(mode, i, s0, ns, wm, "OK" if np.array_equal(got, ref) else f"MISMATCH {np.abs(got-ref).max()}", flush=True)
                except Exception as e:
                    print(mode, i, s0, ns, wm, "EXC", str(e)[:100], flush=True)
                    sys.exit(1)
