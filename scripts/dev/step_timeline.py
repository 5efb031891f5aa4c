"""Timeline of ONE graph replay of the cfg 3 training step (kineto / CUPTI): every kernel with its start offset,
duration and the idle gap in front of it - where does the step's wall time go beyond the sum of its kernels?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
N = 4096
ro, rd = bench.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
with torch.no_grad():
    model.sigma_out.bias.fill_(0.3)
opt = FusedAdam(model.parameters(), lr=5e-4)
st = pipeline.GraphedTrainStep(model, opt, bands, N, 2.0, 6.0, 64, 128)
st(ro, rd, target)
for _ in range(20):
    st.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
n = len(ev) // 3
ev = ev[2 * n:]                      # the last replay
t0 = ev[0].time_range.start
end_prev = t0
tot_k, tot_gap = 0.0, 0.0
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    gap = s - end_prev
    print("%9.1f us  +%7.1f us  gap %6.1f  %s" % (s - t0, t - s, gap, e.name[:100]))
    tot_k += t - s
    if gap > 0:
        tot_gap += gap
    end_prev = max(end_prev, t)
print("replay: %.1f us wall, kernels %.1f us (sum), idle gaps %.1f us, %d kernels" % (end_prev - t0, tot_k, tot_gap, len(ev)))
