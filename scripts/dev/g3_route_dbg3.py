import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_mlp import NeRFWithDINO
from nfs_b200 import pipeline, ops, mlp_g3, _lib
from nfs_b200._lib import ptr
from nfs_b200.mlp import freqs_on, bands_are_octaves
from oracle import nerf_oracle as O
cuda = torch.device("cuda:0")
N, S = 300, 48
ro, rd = O.lego_rays(N, H=128, W=128, seed=3)
ro, rd = ro.to(cuda), rd.to(cuda)
torch.manual_seed(8)
mod = NeRFWithDINO(pos_freq=12, dino_dim=64)
with torch.no_grad():
    mod.density_mlp.density_head.bias.fill_(0.3)
mod = mod.to(cuda)
fmap = torch.randn(1, 9, 9, 64, generator=torch.Generator().manual_seed(1)).to(cuda)
pose = torch.eye(4); pose[2, 3] = 4.0
pinv = torch.inverse(pose).contiguous().to(cuda)
focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
t_rand = torch.rand(N, S, generator=torch.Generator().manual_seed(2)).to(cuda)
pts, z = ops.sample_stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
x = pts.reshape(-1, 3).contiguous()
d = rd.unsqueeze(1).expand(-1, S, -1).reshape(-1, 3).contiguous()
plan = mod._get_plan()
plan.refresh()
bands = mod.pos_encoder.freq_bands
names = "c16 sa sa_bits gate c2 sb density cat16 k1 k2 rgb sb_bits".split()
P = x.shape[0]
runs = []
for it in range(3):
    fr = freqs_on(x.device, bands)
    c16 = torch.empty((P, plan.k0), device=cuda, dtype=torch.bfloat16)
    _lib.call("nfs_g3_operand", ptr(x), ptr(pinv), float(focal), 128, 128, ptr(fmap[0]), 9, 9, 64, ptr(fr), int(fr.numel()),
              int(bands_are_octaves(bands)), P, plan.k0, c16.stride(0), ptr(c16), ops._stream())
    with torch.no_grad():
        rgb, den, saved = plan.run_forward(None, d, None, bands, mod.dir_encoder.freq_bands, c16=c16)
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
