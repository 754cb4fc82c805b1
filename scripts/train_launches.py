"""One eager tensor-core training step (batch 8 of 256^2) between cudaProfilerStart / Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
(per-launch list of the step: which kernels and which layer shapes hold the time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmu_b200
from pmu_b200.synthetic import phantom_volume, phantom_labels

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
trainer = pmu_b200.ProbUNetTrainer("cuda", n_channels=1, n_classes=3, latent_dim=6, beta=10, precision="bf16")
net = trainer.net.train()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
vol, lab = phantom_volume(256, seed=3), phantom_labels(256)
mx = vol[40:40 + B].amax(dim=(1, 2), keepdim=True)
imgs = (vol[40:40 + B] / mx)[:, None].contiguous().cuda()
masks = lab[40:40 + B, None].contiguous().cuda()
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    pmu_b200.dp_train_step(trainer, imgs, masks, opt, graph=False)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("done")
