"""Per-layer timing of every tcgen05 conv launch of one slice batch of the trainer network
(B slices of 256x256), CUDA events, median of 5.  Usage: python scripts/time_convs.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmu_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# (name, C0, C1, Cout, H, ntaps, pool)
LAYERS = [
    ("inc.c2/prior0.c2 64->64 @256 (+pool)", 64, 0, 64, 256, 9, 0),
    ("d1.c1 64->128 @128", 64, 0, 128, 128, 9, None),
    ("d1.c2 128->128 @128 (+pool)", 128, 0, 128, 128, 9, 0),
    ("d2.c1 128->256 @64", 128, 0, 256, 64, 9, None),
    ("d2.c2 256->256 @64 (+pool)", 256, 0, 256, 64, 9, 0),
    ("d3.c1 256->512 @32", 256, 0, 512, 32, 9, None),
    ("d3.c2 512->512 @32 (+pool)", 512, 0, 512, 32, 9, 0),
    ("d4.c1 512->1024 @16", 512, 0, 1024, 16, 9, None),
    ("d4.c2 1024->1024 @16", 1024, 0, 1024, 16, 9, None),
    ("up1.T 1024->512 @16", 1024, 0, 512, 16, 4, None),
    ("up1.c1 512+512->512 @32", 512, 512, 512, 32, 9, None),
    ("up1.c2 512->512 @32", 512, 0, 512, 32, 9, None),
    ("up2.T 512->256 @32", 512, 0, 256, 32, 4, None),
    ("up2.c1 256+256->256 @64", 256, 256, 256, 64, 9, None),
    ("up2.c2 256->256 @64", 256, 0, 256, 64, 9, None),
    ("up3.T 256->128 @64", 256, 0, 128, 64, 4, None),
    ("up3.c1 128+128->128 @128", 128, 128, 128, 128, 9, None),
    ("up3.c2 128->128 @128", 128, 0, 128, 128, 9, None),
    ("up4.T 128->64 @128", 128, 0, 64, 128, 4, None),
    ("up4.c1 64+64->64 @256", 64, 64, 64, 256, 9, None),
    ("up4.c2 64->64 @256", 64, 0, 64, 256, 9, None),
]
# multiplicity per slice: encoder convs appear in unet AND prior (same shapes)
MULT = {0: 2, 1: 2, 2: 2, 3: 2, 4: 2, 5: 2, 6: 2, 7: 2, 8: 2}
tot_ms, tot_fl = 0.0, 0.0
print(f"variant={os.environ.get('PMU_CONV_VARIANT', '0')} B={B}")
for i, (name, C0, C1, Cout, H, ntaps, pool) in enumerate(LAYERS):
    x0 = torch.randn(B, H, H, C0, device="cuda").to(torch.float16)
    x1 = torch.randn(B, H, H, C1, device="cuda").to(torch.float16) if C1 else None
    K = (9 if ntaps == 9 else 1) * (C0 + C1)
    N = (4 if ntaps == 4 else 1) * Cout
    wp = (torch.randn(N, K, device="cuda") * 0.01).to(torch.float16)
    bias = torch.zeros(Cout, device="cuda")
    def run():
        if pool is not None:
            ops.conv_gemm_pool_bf16(x0, wp, bias, Cout, True, pool)
        else:
            ops.conv_gemm_bf16(x0, wp, bias, Cout, ntaps, ntaps != 4, x1=x1)
    for _ in range(2): run()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    fl = 2.0 * B * H * H * N * K
    m = MULT.get(i, 1)
    tot_ms += m * ms; tot_fl += m * fl
    print(f"{name:40s} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s  x{m}")
print(f"TOTAL per batch of {B}: {tot_ms:.2f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s; per 768 slices: {tot_ms * 768 / B:.1f} ms")
