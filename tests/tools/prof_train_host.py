"""Where does the wall time of a bf16 training step go?  Phase timings with synchronisation + torch profiler top ops."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, pmu_b200
from oracle import pmu_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
trainer = pmu_b200.ProbUNetTrainer("cuda", n_channels=1, n_classes=3, latent_dim=6, beta=10, precision="bf16")
net = trainer.net.train()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
vol, lab = O.phantom(256, seed=3)
imgs = torch.from_numpy(O.plane_slices(vol, 0, 100, B)).cuda()
masks = torch.from_numpy(lab[100:100 + B, None].astype(np.float32)).cuda()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for it in range(NIT):
    t0 = T(); net.forward(imgs, masks, training=True)
    t1 = T(); s = net.sample(testing=False)
    t2 = T(); loss = -net.elbo(masks)
    t3 = T(); loss.backward()
    t4 = T(); torch.nn.utils.clip_grad_value_(net.parameters(), 0.1)
    t5 = T(); opt.step(); opt.zero_grad()
    t6 = T()
    print(f"it {it}: forward {1e3*(t1-t0):.1f}  sample {1e3*(t2-t1):.1f}  elbo {1e3*(t3-t2):.1f}  backward {1e3*(t4-t3):.1f}  clip {1e3*(t5-t4):.1f}  sgd {1e3*(t6-t5):.1f}  total {1e3*(t6-t0):.1f} ms")
print("alloc retries:", torch.cuda.memory_stats().get("num_alloc_retries"), " peak GB:", torch.cuda.max_memory_allocated() / 1e9)
if '--no-profiler' in sys.argv:
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.predict(imgs, masks); loss = trainer.loss(imgs, masks, None); loss.backward(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=50))
