"""Run the TMA-brick resampling gather alone on an oblique grid (for ncu / timing): 256^3 volume, 256 slices of 256 x 256."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmu_b200
from pmu_b200 import ops
from pmu_b200.synthetic import phantom_volume
D = 256
vol = phantom_volume(D).cuda()
aff = [0.5, -0.25, 0.75, 1.0, 0.01, 0.0, -0.01, 1.0, 0.02, 0.0, -0.02, 1.0]
out = torch.empty(D, 1, D, D, device="cuda")
for mode in ("trilinear", "nearest"):
    for _ in range(2):
        ops.slice_gather(vol, 0, 0, D, interp=mode, affine=aff, hw=(D, D), out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.slice_gather(vol, 0, 0, D, interp=mode, affine=aff, hw=(D, D), out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{mode:9s} {ms * 1e3:7.1f} us  {D ** 3 * 8 / ms / 1e6:7.1f} GB/s  frac {D ** 3 * 8 / ms / 1e6 / 6553.9:.3f}")
