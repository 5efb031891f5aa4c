"""Per-CTA start / end times of the merged backward kernel (NFS_BWD_TIMES): who finishes when - the dgrad producers or
the consumers of which weight-gradient job.  Usage: NFS_BWD_PRODUCERS=46 python scripts/dev/bwd_times.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
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
    model = NeRFMLP().to(dev).train()
    with torch.no_grad():
        model.sigma_out.bias.fill_(0.3)
    opt = FusedAdam(model.parameters(), lr=5e-4)
    for i in range(4):
        print("bwtimes step %d" % i, flush=True)
        pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)
        torch.cuda.synchronize()
    sys.exit(0)
env = dict(os.environ, NFS_BWD_MERGED="1", NFS_BWD_TIMES="1")
out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("bwtimes")]
last = max(i for i, l in enumerate(lines) if l.startswith("bwtimes step"))
rows = lines[last + 1:]
if not rows:
    print(out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(1)
prod, cons = [], {}
for l in rows:
    w = l.split()
    if w[1] == "producer":
        prod.append((int(w[5]), int(w[7])))
    else:
        cons.setdefault(int(w[5]), []).append((int(w[9]), int(w[11]), int(w[7])))
t0 = min(s for s, e in prod)
print("producers: %d CTAs, end %.1f .. %.1f us" % (len(prod), min(e - t0 for s, e in prod) / 1e3, max(e - t0 for s, e in prod) / 1e3))
for j in sorted(cons):
    c = cons[j]
    print("job %d: %d CTAs, end %.1f .. %.1f us" % (j, len(c), min(e - t0 for s, e, n in c) / 1e3, max(e - t0 for s, e, n in c) / 1e3))
