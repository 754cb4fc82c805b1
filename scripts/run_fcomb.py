"""Run the fused fcomb kernel alone (for ncu / timing): B slices of 256x256, N samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmu_b200
from pmu_b200 import ops
from pmu_b200.synthetic import trainer_state_dict
from pmu_b200.engine import PackedNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sd = trainer_state_dict(0)
fw = PackedNet({k: v for k, v in sd.items() if k.startswith("fcomb")}, "cuda", "f16").fcomb
feat = torch.relu(torch.randn(B, 256, 256, 64, device="cuda")).to(torch.float16)
mu = torch.randn(B, 6, device="cuda"); sigma = torch.rand(B, 6, device="cuda") + 0.2
eps = torch.randn(B, N, 6, device="cuda")
for _ in range(2): ops.fcomb_softmax_accum_bf16(feat, mu, sigma, eps, fw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.fcomb_softmax_accum_bf16(feat, mu, sigma, eps, fw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
fl = B * 65536 * (2 * 64 * 64 + N * (2 * 2 * 64 * 64 + 2 * 64 * 3))
print(f"fcomb B={B} N={N}: {ms:.3f} ms/launch  -> {ms * 768 / B:.1f} ms per 768 slices; {fl / ms / 1e9:.1f} TFLOP/s (useful)")
