"""Timing of the small-N conv layers only (64/128 couts), for bottleneck experiments (PMU_CONV_DEBUG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmu_b200 import ops
B = 64
LAYERS = [("64->64 @256 +pool", 64, 0, 64, 256, 0), ("64->64 @256", 64, 0, 64, 256, None), ("64+64->64 @256", 64, 64, 64, 256, None),
          ("64->128 @128", 64, 0, 128, 128, None), ("128->128 @128", 128, 0, 128, 128, None), ("128+128->128 @128", 128, 128, 128, 128, None)]
print("debug =", os.environ.get("PMU_CONV_DEBUG", "0"), " rs =", os.environ.get("PMU_CONV_RS", "1"))
for name, C0, C1, Cout, H, pool in LAYERS:
    x0 = torch.randn(B, H, H, C0, device="cuda").to(torch.float16)
    x1 = torch.randn(B, H, H, C1, device="cuda").to(torch.float16) if C1 else None
    wp = (torch.randn(Cout, 9 * (C0 + C1), device="cuda") * 0.01).to(torch.float16)
    bias = torch.zeros(Cout, device="cuda")
    def run():
        if pool is not None: ops.conv_gemm_pool_bf16(x0, wp, bias, Cout, True, pool)
        else: ops.conv_gemm_bf16(x0, wp, bias, Cout, 9, True, x1=x1)
    for _ in range(2): run()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"{name:22s} {ms:7.3f} ms {2.0 * B * H * H * Cout * 9 * (C0 + C1) / ms / 1e9:7.1f} TFLOP/s")
