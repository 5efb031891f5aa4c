"""First-call flake of smoke()'s compositing check (1 in ~10 fresh processes: a 128-ray block off by ~1e-4): which side
varies - the CUDA kernel or ATen's CPU arithmetic of the oracle?  Prints a hash of both for two evaluations each."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
n, S = 2048, 64
g = torch.Generator().manual_seed(0)
ro, rd = O.lego_rays(n)
t_rand = torch.rand(n, S, generator=g)
rgb = torch.rand(n, S, 3, generator=g)
den = torch.randn(n, S, 1, generator=g) * 10
pts, z = ops.sample_stratified(ro.to(dev), rd.to(dev), 2.0, 6.0, S, t_rand=t_rand.to(dev))
pts_o, z_o = O.stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
h = lambda t: hashlib.md5(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:8]
gpu, cpu = [], []
for t in range(2):
    rgb_d, den_d = rgb.to(dev).requires_grad_(), den.to(dev).requires_grad_()
    gpu.append(ops.composite(rgb_d, den_d, z, rd.to(dev))[0].detach().cpu())
    rgb_c, den_c = rgb.clone().requires_grad_(), den.clone().requires_grad_()
    cpu.append(O.render(rgb_c, den_c, z_o, rd)[0].detach())
d = (cpu[0] - cpu[1]).abs().max(1).values
bad = (d > 0).nonzero().flatten().tolist()
print("threads", torch.get_num_threads(), "gpu", h(gpu[0]), h(gpu[1]), "cpu", h(cpu[0]), h(cpu[1]),
      "cpu call0 vs call1: max", float(d.max()), "rays", (bad[0], bad[-1], len(bad)) if bad else None)
