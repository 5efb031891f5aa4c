import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_mlp import NeRFWithDINO
from nfs_b200 import pipeline
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
pose = pose.to(cuda)
focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
t_rand = torch.rand(N, S, generator=torch.Generator().manual_seed(2)).to(cuda)
res = []
for mode in ("1", "0", "1", "0"):
    os.environ["NFS_G3_OPERAND"] = mode
    with torch.no_grad():
        out = pipeline.render_rays_conditioned(mod, ro, rd, 2.0, 6.0, S, pose, focal, 128, 128, fmap, perturb=True, t_rand=t_rand)
    res.append(out["rgb"].clone())
for i in range(4):
    for j in range(i + 1, 4):
        d = (res[i] - res[j]).abs()
        print(i, j, "max diff", float(d.max()), "n diff", int((d > 0).sum()))
# operand comparison on the pipeline's own points
from nfs_b200 import ops, mlp_g3, _lib
from nfs_b200._lib import ptr
from nfs_b200.mlp import encode_operand
pts, z = ops.sample_stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
x = pts.reshape(-1, 3)
pinv = torch.inverse(pose).contiguous()
_, _, _, feats = ops.project_gather(x, pose, focal, 128, 128, features=fmap, want_projection=False, pose_inv=pinv)
plan = mod._get_plan()
ref = encode_operand(x, mod.pos_encoder.freq_bands, plan.k0, extra=feats)
out = torch.empty_like(ref)
fr = mod.pos_encoder.freq_bands.to(cuda).float().contiguous()
_lib.call("nfs_g3_operand", ptr(x), ptr(pinv), float(focal), 128, 128, ptr(fmap), 9, 9, 64, ptr(fr), 12, 1, x.shape[0], plan.k0, plan.k0, ptr(out), ops._stream())
d = (out.float() - ref.float()).abs()
rows, cols = (d > 0).nonzero(as_tuple=True)
print("operand: differing elements", len(rows), "cols", sorted(set(cols.tolist()))[:20], "max", float(d.max()))
