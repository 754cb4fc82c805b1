"""Diagnostics for the tcgen05 conv kernel (run on the GPU box): per-tap / per-shape error report
against torch fp32 CPU conv on bf16-rounded operands."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import pmu_b200
from pmu_b200 import ops

def bf(t): return t.to(torch.bfloat16).float()
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()

def run(B, C0, C1, Cout, H, W, tap_only=None, relu=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    x0 = bf(torch.randn(B, C0, H, W, generator=g)); x1 = bf(torch.randn(B, C1, H, W, generator=g)) if C1 else None
    Cin = C0 + C1
    w = bf(torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5)
    if tap_only is not None:
        m = torch.zeros(3, 3); m.view(-1)[tap_only] = 1; w = w * m
    b = torch.randn(Cout, generator=g) * 0.1
    xin = torch.cat([x0, x1], 1) if C1 else x0
    ref = F.conv2d(xin, w, b, padding=1)
    if relu: ref = F.relu(ref)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(torch.bfloat16).contiguous().cuda()
    got = ops.conv_gemm_bf16(nhwc(x0).to(torch.bfloat16).cuda(), wp, b.cuda(), Cout, 9, relu,
                             nhwc(x1).to(torch.bfloat16).cuda() if C1 else None)
    torch.cuda.synchronize()
    got = got.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    return float(err.max()), float(ref.abs().max()), got, ref

if __name__ == "__main__":
    print("device", torch.cuda.get_device_name(0))
    for tap in [None] + list(range(9)):
        e, m, got, ref = run(1, 64, 0, 64, 16, 16, tap_only=tap, relu=False)
        print(f"64->64 16x16 tap={tap}: max_err={e:.4f} ref_max={m:.3f}")
        if tap is None and e > 0.05:
            # where are the errors? per-row / per-channel summary
            d = (got - ref).abs()[0]
            print(" err by channel block of 8:", [round(float(d[c:c+8].max()), 3) for c in range(0, 64, 8)])
            print(" err by row:", [round(float(d[:, r].max()), 3) for r in range(16)])
            print(" err by col:", [round(float(d[:, :, c].max()), 3) for c in range(16)])
            print(" got[0,:4,0,:4]", got[0, :4, 0, :4]); print(" ref[0,:4,0,:4]", ref[0, :4, 0, :4])
    for case in [(2, 64, 0, 128, 32, 32), (3, 128, 128, 128, 16, 16), (5, 256, 0, 256, 4, 4), (3, 64, 0, 64, 2, 2),
                 (2, 64, 64, 64, 24, 40), (1, 512, 0, 1024, 8, 8), (4, 64, 0, 64, 256, 256)]:
        e, m, _, _ = run(*case)
        print(f"case {case}: max_err={e:.4f} ref_max={m:.3f}")
    # timing of a big layer
    for (B, C0, Cout, H) in [(64, 64, 64, 256), (64, 128, 128, 128), (64, 256, 256, 64), (64, 512, 512, 32), (64, 1024, 1024, 16)]:
        x = torch.randn(B, H, H, C0, device="cuda").to(torch.bfloat16)
        wp = (torch.randn(Cout, 9 * C0, device="cuda") * 0.01).to(torch.bfloat16)
        bias = torch.zeros(Cout, device="cuda")
        for _ in range(2): ops.conv_gemm_bf16(x, wp, bias, Cout, 9, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.conv_gemm_bf16(x, wp, bias, Cout, 9, True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * H * H * Cout * 9 * C0
        print(f"conv {C0}->{Cout} @{H}^2 x{B}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
