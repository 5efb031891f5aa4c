"""A/B timing of the graph-replayed training step (cfg 3) under environment switches, same process / same box.
usage: ab_step.py NAME=VALUE[,NAME=VALUE...] ...   (each argument is one variant; "-" = defaults)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
variants = sys.argv[1:] or ["-"]
N = 4096
ro, rd = bench.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
steps = []
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",")) if v != "-" else {}
    os.environ.update(env)
    torch.manual_seed(0)
    model = NeRFMLP().to(dev).train()
    opt = FusedAdam(model.parameters(), lr=5e-4)
    steps.append(pipeline.GraphedTrainStep(model, opt, bands, N, 2.0, 6.0, 64, 128))
    for k in env:
        del os.environ[k]
for rnd in range(3):
    for v, st in zip(variants, steps):
        for _ in range(5):
            st(ro, rd, target)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(30):
            st(ro, rd, target)
        b.record(); torch.cuda.synchronize()
        print("round %d  %-40s %.3f ms/step" % (rnd, v, a.elapsed_time(b) / 30))
