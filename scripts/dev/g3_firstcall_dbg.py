import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_mlp import NeRFWithDINO
cuda = torch.device("cuda:0")
torch.manual_seed(8)
mod = NeRFWithDINO(pos_freq=12, dino_dim=64).to(cuda)
P = 14400
g = torch.Generator().manual_seed(1)
x = ((torch.rand(P, 3, generator=g) - 0.5) * 6).to(cuda)
d = torch.randn(P, 3, generator=g).to(cuda)
f = torch.randn(P, 64, generator=g).to(cuda)
plan = mod._get_plan()
names = "c16 sa sa_bits gate c2 sb density cat16 k1 k2 rgb sb_bits".split()
runs = []
for i in range(3):
    plan.refresh()
    with torch.no_grad():
        rgb, den, saved = plan.run_forward(x, d, f, mod.pos_encoder.freq_bands, mod.dir_encoder.freq_bands)
    torch.cuda.synchronize()
    runs.append([t.clone() for t in saved])
for i in (1, 2):
    print("run 0 vs run", i)
    for n, a, b in zip(names, runs[0], runs[i]):
        if a.dim() == 3:
            a, b = a[:, :P], b[:, :P]
        if n == "sa":
            a, b = torch.cat([a[:2].flatten(), a[2, :, :128].flatten()]), torch.cat([b[:2].flatten(), b[2, :, :128].flatten()])
        neq = int((a != b).sum())
        print("   %-8s differing %d  max %g" % (n, neq, float((a.float() - b.float()).abs().max()) if neq else 0.0))
