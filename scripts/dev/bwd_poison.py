"""Early-read detector for the merged backward kernel: the step is run once (reference gradient), then the dY arena is
filled with NaN and the same step (same seeds) is run again.  A weight-gradient consumer that loads a row before the
dgrad chain has stored it picks up NaN - which parameter tensors are affected says which hand-off is broken.
usage: [NFS_K1_BWD_DY=0|1] python scripts/dev/bwd_poison.py [n_rays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
import bench
from nfs_b200 import pipeline
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 700
ro, rd = bench.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)
bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
with torch.no_grad():
    model.sigma_out.bias.fill_(0.3)
opt = FusedAdam(model.parameters(), lr=0.0)
names = [n for n, _ in model.named_parameters()]
sizes = [p.numel() for p in model.parameters()]
ref = None
for rep in range(6):
    sess = pipeline._session_for(model, opt)
    if rep > 0:
        sess.dys.fill_(float("nan"))
        if os.environ.get("POISON_DY", "0") != "0":
            sess.dy.fill_(float("nan"))
    torch.manual_seed(11)
    pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 48, 80)
    g = opt.grad.clone()
    if ref is None:
        ref = g
        print("merged:", sess.merged, " rows:", sess.cursor, flush=True)
        continue
    bad, off = [], 0
    for n, s in zip(names, sizes):
        part, rpart = g[off:off + s], ref[off:off + s]
        nn = int(torch.isnan(part).sum())
        rel = float((torch.nan_to_num(part) - rpart).norm() / rpart.norm().clamp_min(1e-20))
        if nn or rel > 1e-5:
            bad.append("%s: %d NaN, rel %.1e" % (n, nn, rel))
        off += s
    print("rep %d: %s" % (rep, "; ".join(bad) if bad else "clean"), flush=True)
