"""Run-to-run determinism of the training step's gradient (merged backward kernel): the same step five times from the
same state; differences beyond the weight-gradient kernels' fp32 atomics (1e-6) mean a hand-off race."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
for N in (700, 4096):
    ro, rd = bench.lego_rays(N, seed=0)
    ro, rd = ro.to(dev), rd.to(dev)
    target = torch.rand(N, 3, device=dev)
    bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
    gs = []
    for rep in range(5):
        torch.manual_seed(0)
        model = NeRFMLP().to(dev).train()
        with torch.no_grad():
            model.sigma_out.bias.fill_(0.3)
        opt = FusedAdam(model.parameters(), lr=5e-4)
        torch.manual_seed(11)
        pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)
        gs.append(opt.grad.clone())
    print("N=%d  rel. L2 of runs 2..5 against run 1:" % N, " ".join("%.2e" % float((g - gs[0]).norm() / gs[0].norm()) for g in gs[1:]), flush=True)
