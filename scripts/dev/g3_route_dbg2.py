import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_mlp import NeRFWithDINO
from nfs_b200 import pipeline, ops, mlp_g3
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
runs = []
for it in range(3):
    with torch.no_grad():
        pts, z = ops.sample_stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
        pts_flat = pts.reshape(-1, 3)
        dirs = rd.unsqueeze(1).expand(-1, S, -1).reshape(-1, 3)
        plan = mod._get_plan()
        pinv = torch.inverse(pose)
        rgb, density = mlp_g3.g3_forward_from_map(plan, pts_flat, dirs, fmap, pinv, focal, 128, 128)
        rgb_map, depth, weights = ops.composite(rgb.reshape(N, S, 3), density.reshape(N, S, 1), z, rd, white_bkgd=False)
    torch.cuda.synchronize()
    runs.append(dict(pts=pts.clone(), z=z.clone(), pinv=pinv.clone(), rgb=rgb.clone(), density=density.clone(), rgb_map=rgb_map.clone(),
                     depth=depth.clone(), weights=weights.clone()))
for i in (1, 2):
    print("run 0 vs", i, {k: (int((runs[0][k] != runs[i][k]).sum()), float((runs[0][k] - runs[i][k]).abs().max())) for k in runs[0]})

print("---- module API route, three calls")
runs = []
for it in range(3):
    with torch.no_grad():
        _, _, _, feats = ops.project_gather(pts_flat, pose, focal, 128, 128, features=fmap, want_projection=False, pose_inv=pinv.contiguous())
        rgb, density = mod(pts_flat, dirs, feats)
    torch.cuda.synchronize()
    runs.append(dict(rgb=rgb.clone(), density=density.clone()))
for i in (1, 2):
    print("run 0 vs", i, {k: (int((runs[0][k] != runs[i][k]).sum()), float((runs[0][k] - runs[i][k]).abs().max())) for k in runs[0]})
print("---- fresh model, operand route, explicit refresh first")
torch.manual_seed(8)
mod2 = NeRFWithDINO(pos_freq=12, dino_dim=64)
with torch.no_grad():
    mod2.density_mlp.density_head.bias.fill_(0.3)
mod2 = mod2.to(cuda)
plan2 = mod2._get_plan()
plan2.refresh()
torch.cuda.synchronize()
runs = []
for it in range(3):
    with torch.no_grad():
        rgb, density = mlp_g3.g3_forward_from_map(plan2, pts_flat, dirs, fmap, pinv, focal, 128, 128)
    torch.cuda.synchronize()
    runs.append(dict(rgb=rgb.clone(), density=density.clone()))
for i in (1, 2):
    print("run 0 vs", i, {k: (int((runs[0][k] != runs[i][k]).sum()), float((runs[0][k] - runs[i][k]).abs().max())) for k in runs[0]})
