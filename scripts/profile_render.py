"""Per-kernel breakdown of the full-frame render path (BASELINE config 5) for one 65 536-ray chunk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from nfs_b200 import pipeline
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev).eval()
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
N = 65536 * 2
ro, rd = bench.lego_rays(N, seed=7)
ro, rd = ro.to(dev), rd.to(dev)
f = lambda: pipeline.render_image(model, bands, ro, rd, 2.0, 6.0, 64, 128, chunk=65536)
for _ in range(3):
    f()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    f()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA":
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("GPU kernel time for %d rays: %.3f ms over %d launches" % (N, tot / 1e3, sum(r[1] for r in rows)))
for t, n, k in rows[:14]:
    print("%9.1f us  %4d x  %5.1f%%  %s" % (t, n, 100 * t / tot, k[:100]))
