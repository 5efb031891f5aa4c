"""Three BASELINE-config-3 train steps (4096 rays, 64 + 192 evaluations per ray) - the command the
per-launch ncu list of the train step is taken from."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from oracle import nerf_oracle as O
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
opt = FusedAdam(model.parameters(), lr=5e-4)
bands = O.frequency_bands(10)
N = 4096
ro, rd = O.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loss = pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)
torch.cuda.synchronize()
print("loss", float(loss))
