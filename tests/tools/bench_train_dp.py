"""Data-parallel training step (BASELINE config 4): global batch 8*world of 256x256 slices, bf16 tensor-core
mode, one process per GPU, gradients SUM-all-reduced over NCCL in buckets (pmu_b200.train_dp).  Launch with
torchrun.  Prints ms/step (max over ranks, CUDA events) and checks that all ranks hold identical weights after."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
import pmu_b200
from oracle import pmu_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 8
torch.manual_seed(0)
trainer = pmu_b200.ProbUNetTrainer(f"cuda:{local}", n_channels=1, n_classes=3, latent_dim=6, beta=10, precision="bf16")
net = trainer.net.train()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
vol, lab = O.phantom(256, seed=3)
s0 = 40 + rank * B
imgs = torch.from_numpy(O.plane_slices(vol, 0, s0, B)).cuda()
masks = torch.from_numpy(lab[s0:s0 + B, None].astype(np.float32)).cuda()
steps = 8
for _ in range(6):     # the caching allocator still grows during the first five steps (one late cudaMalloc)
    pmu_b200.dp_train_step(trainer, imgs, masks, opt)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = pmu_b200.dp_train_step(trainer, imgs, masks, opt)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
gl = loss.clone(); dist.all_reduce(gl)
# replicas must stay bit-identical: same initial weights + identical summed gradients
chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"DP train step: world {world}, global batch {B * world} x 256x256 bf16: {float(ms):.1f} ms/step, "
          f"{B * world / float(ms) * 1e3:.1f} slices/s, global loss {float(gl):.1f}, replicas identical: {bool(lo == hi)}")
dist.destroy_process_group()
