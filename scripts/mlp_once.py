"""Fused MLP forward (inference, 786432 points = 4096 rays x 192 samples) and one training
forward+backward - the command the ncu capture of the K3 kernels is taken from."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from oracle import nerf_oracle as O
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev)
bands = O.frequency_bands(10)
P = 4096 * 192
pts = (torch.rand(P, 3, device=dev) - 0.5) * 6
with torch.no_grad():
    for _ in range(3):
        out = model.forward_points(pts, bands)
for _ in range(2):
    out = model.forward_points(pts, bands)
    torch.autograd.grad(out, list(model.parameters()), torch.ones_like(out))
torch.cuda.synchronize()
print("ok", float(out.sum()))
