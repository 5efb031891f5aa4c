"""Per-kernel breakdown of one training step of BASELINE config 4 (NeRFWithDINO, 512 rays x 64 samples) with
torch.profiler; same set-up as bench.py's dino_nerf_cfg4."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_mlp import NeRFWithDINO

dev = torch.device("cuda:0")
torch.manual_seed(1)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g3 = NeRFWithDINO(pos_freq=12, dir_freq=4, dino_dim=64).to(dev).train()
opt = FusedAdam(g3.parameters(), lr=5e-4)
ro, rd = bench.lego_rays(nb, H=128, W=128, seed=200)
ro, rd = ro.to(dev), rd.to(dev)
tgt = torch.rand(nb, 3, device=dev)
fmap = torch.randn(1, 9, 9, 64, device=dev)
pose = torch.eye(4, device=dev); pose[2, 3] = 4.0
focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
pose_inv = torch.inverse(pose)

def step():
    opt.zero_grad()
    o = pipeline.render_rays_conditioned(g3, ro, rd, 2.0, 6.0, 64, pose, focal, 128, 128, fmap, perturb=True, pose_inv=pose_inv)
    loss = torch.mean((o["rgb"] - tgt) ** 2)
    loss.backward()
    opt.gather_grads()
    opt.step(gathered=True)

for _ in range(5):
    step()
torch.cuda.synchronize()
reps = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA":
        rows.append((t / reps, e.count / reps, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("GPU kernel time per step: %.3f ms over %.1f launches" % (tot / 1e3, sum(r[1] for r in rows)))
for t, n, k in rows[:32]:
    print("%9.1f us  %5.1f x  %s" % (t, n, k[:120]))
