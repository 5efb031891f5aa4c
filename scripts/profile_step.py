"""Per-kernel breakdown of one training step of BASELINE config 3 (4096 rays, 64 + 192 evaluations
per ray, G1 model) with torch.profiler (kineto / CUPTI; no nsys in this image)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
opt = FusedAdam(model.parameters(), lr=5e-4)
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
N = 4096
ro, rd = bench.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)
step = lambda: pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)
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
for t, n, k in rows[:40]:
    print("%9.1f us  %5.1f x  %s" % (t, n, k[:110]))
