"""Training-step timing (BASELINE config 4 shape, fp32 parity mode): B slices of 256x256 through
ProbUNetTrainer.predict -> loss -> backward (+ clip + SGD step), CUDA events, plus the per-entry-point
breakdown and the oracle (torch CPU autograd) on a smaller batch for the same step.
usage: python tests/tools/bench_train.py [B] [steps] [--cpu]"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import pmu_b200
from pmu_b200 import ops
from oracle import pmu_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
torch.manual_seed(0)
PREC = "bf16" if "--bf16" in sys.argv else "fp32"
trainer = pmu_b200.ProbUNetTrainer("cuda", n_channels=1, n_classes=3, latent_dim=6, beta=10, precision=PREC)
net = trainer.net.train()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
vol, lab = O.phantom(256, seed=3)
imgs = torch.from_numpy(O.plane_slices(vol, 0, 100, B)).cuda()
masks = torch.from_numpy(lab[100:100 + B, None].astype(np.float32)).cuda()

def step():
    trainer.predict(imgs, masks)
    loss = trainer.loss(imgs, masks, None)
    loss.backward()
    torch.nn.utils.clip_grad_value_(net.parameters(), 0.1)
    opt.step(); opt.zero_grad()
    return loss

for _ in range(6): step()          # warm-up (the caching allocator still issues one cudaMalloc at step 5): module load, allocator, lazy cudaFuncSetAttribute (iteration 1 is still 8x slower)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
import gc
if "--gc-off" in sys.argv:
    gc.disable()
mallocs, host = [], []
evs[0].record()
for i in range(steps):
    m0, t0 = torch.cuda.memory_stats().get("num_device_alloc", 0), time.perf_counter()
    loss = step()
    evs[i + 1].record()
    mallocs.append(torch.cuda.memory_stats().get("num_device_alloc", 0) - m0)
    host.append(1e3 * (time.perf_counter() - t0))
torch.cuda.synchronize()
print("cudaMalloc calls per step:", mallocs, "| host ms per step:", " ".join(f"{t:.1f}" for t in host),
      "| reserved GB", round(torch.cuda.memory_reserved() / 1e9, 2))
per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
ms = evs[0].elapsed_time(evs[steps]) / steps
print("per-step ms:", " ".join(f"{t:.1f}" for t in per), f"| median {sorted(per)[len(per) // 2]:.1f}  mean {ms:.1f}")
# FLOPs: forward 96.2 (unet) + 2 x 33.9 (prior, posterior) + 2 x 1.69 (fcomb: sample + reconstruction) GFLOP per 256^2 slice;
# backward = dgrad + wgrad ~ 2 x forward of the differentiated part
fwd = 96.18 + 2 * 33.90 + 2 * 1.69
flop = B * (fwd + 2 * (fwd - 1.69)) * 1e9
print(f"train step B={B} x 256x256 {PREC}: {ms:.1f} ms/step, {B / ms * 1e3:.2f} slices/s, {flop / ms / 1e9:.1f} TFLOP/s (useful), loss {float(loss):.1f}")
ops.PROFILE = []
step(); torch.cuda.synchronize()
tot = collections.defaultdict(float)
for name, meta, a, b in ops.PROFILE: tot[name] += a.elapsed_time(b)
ops.PROFILE = None
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]: print(f"  {v:9.2f} ms {100 * v / s:5.1f}%  {k}")
if "--cpu" in sys.argv:
    Bc = 2
    sd = O.make_state_dict(seed=0)
    ps = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and "running_" not in k}
    full = dict(sd); full.update(ps)
    torch.set_num_threads(os.cpu_count())
    x, m = imgs[:Bc].cpu(), masks[:Bc].cpu()
    t0 = time.perf_counter()
    r = O.elbo(full, x, m, torch.randn(Bc, 6), beta=10.0, bn_train=True)
    (-r["elbo"]).backward()
    dt = time.perf_counter() - t0
    print(f"oracle (torch CPU autograd, {os.cpu_count()} threads) B={Bc}: {dt * 1e3:.0f} ms/step, {Bc / dt:.3f} slices/s")
