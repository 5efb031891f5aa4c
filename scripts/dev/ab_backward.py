"""A/B of the backward pass of the cfg 3 step: merged launch (nfs_mlp_backward_fused) at several producer / consumer
splits against the split route (dgrad chain per call + one weight-gradient launch per layer).  Times the eager
backward section with CUDA events (forward excluded) and the whole graph-replayed step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
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
variants = sys.argv[1:] or ["NFS_BWD_MERGED=0", "NFS_BWD_PRODUCERS=32", "NFS_BWD_PRODUCERS=37", "NFS_BWD_PRODUCERS=42",
                            "NFS_BWD_PRODUCERS=46", "NFS_BWD_PRODUCERS=50"]
steps = []
g_ref = None
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",")) if v != "-" else {}
    os.environ.update(env)
    torch.manual_seed(0)
    model = NeRFMLP().to(dev).train()
    with torch.no_grad():
        model.sigma_out.bias.fill_(0.3)          # keeps the random-init density alive (a dead density has no gradient)
    opt = FusedAdam(model.parameters(), lr=5e-4)
    # gradient check: one eager step with fixed draws, against the first variant's gradient
    torch.manual_seed(11)
    pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)
    g = opt.grad.clone()
    if g_ref is None:
        g_ref = g
    print("%-28s |g| %.4e  rel. L2 vs first variant %.3e" % (v, float(g.norm()), float((g - g_ref).norm() / g_ref.norm())), flush=True)
    steps.append((pipeline.GraphedTrainStep(model, opt, bands, N, 2.0, 6.0, 64, 128), opt))
    for k in env:
        del os.environ[k]
ref = None
for v, (st, opt) in zip(variants, steps):
    st(ro, rd, target)
    torch.cuda.synchronize()
for rnd in range(2):
    for v, (st, opt) in zip(variants, steps):
        for _ in range(5):
            st.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(30):
            st.replay()
        b.record(); torch.cuda.synchronize()
        print("round %d  %-28s %.3f ms/step   loss %.5f" % (rnd, v, a.elapsed_time(b) / 30, float(st.loss)), flush=True)
