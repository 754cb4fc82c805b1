"""Per-shape timing of the bf16 NHWC BatchNorm forward / backward of the tensor-core training step (batch 8 of 256^2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmu_b200
from pmu_b200 import ops
tot_f = tot_b = 0.0
for (H, C, n_layers) in [(256, 64, 8), (128, 128, 8), (64, 256, 8), (32, 512, 8), (16, 1024, 6)]:
    y = torch.randn(8, H, H, C, device="cuda").to(torch.bfloat16)
    da = torch.randn(8, H, H, C, device="cuda").to(torch.bfloat16)
    g, b = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    def fwd(): return ops.bn_train_fwd_nhwc_bf16(y, g, b, 1e-5, True, 0.1, rm, rv)
    a, mean, var = fwd()
    def bwd(): return ops.bn_train_bwd_nhwc_bf16(da, y, mean, var, g, b, 1e-5, True)
    res = []
    for fn, nbytes in ((fwd, 3 * y.numel() * 2), (bwd, 5 * y.numel() * 2)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append((us, nbytes / us / 1e3))
    tot_f += res[0][0] * n_layers; tot_b += res[1][0] * n_layers
    print(f"[8,{H},{H},{C}] {y.numel() * 2 / 1e6:6.1f} MB  fwd {res[0][0]:7.1f} us {res[0][1]:7.1f} GB/s   bwd {res[1][0]:7.1f} us {res[1][1]:7.1f} GB/s")
print(f"all 38 layers of a step: fwd {tot_f / 1e3:.2f} ms, bwd {tot_b / 1e3:.2f} ms")
