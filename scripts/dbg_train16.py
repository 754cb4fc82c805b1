import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmu_b200
from oracle import pmu_oracle as O
sd = O.make_state_dict(seed=0)
g = torch.Generator().manual_seed(30)
HW = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(4, 1, HW, HW, generator=g).cuda()
m = torch.randint(0, 3, (4, 1, HW, HW), generator=g).float().cuda()
eps = torch.randn(4, 6, generator=g).cuda()
grads = {}
for prec in ("fp32", "bf16"):
    net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128, 256, 512, 1024], 6, 4, 10)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train().set_precision(prec)
    net.forward(x, m, training=True)
    e = net.elbo(m, eps=eps)
    (-e).backward()
    grads[prec] = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    print(prec, float(e.detach()), float(net.kl), float(net.reconstruction_loss))
for k, gf in grads["fp32"].items():
    gb = grads["bf16"][k]
    rel = float((gb - gf).norm() / gf.norm().clamp_min(1e-12))
    print(f"{rel:8.4f} {float(gf.norm()):10.3e} {k}")
