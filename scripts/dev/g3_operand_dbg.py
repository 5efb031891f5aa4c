import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import _lib, ops
from nfs_b200._lib import ptr
from nfs_b200.mlp import encode_operand, pad_in
from nfs_b200.ops import _stream
cuda = torch.device("cuda:0")
P, pos_freq, C = 5, 12, 64
g = torch.Generator().manual_seed(P)
x = ((torch.rand(P, 3, generator=g) - 0.5) * 8).to(cuda)
fmap = torch.randn(1, 9, 9, C, generator=g).to(cuda)
pose = torch.eye(4); pose[2, 3] = 4.0; pose[0, 3] = 0.3
pose_inv = torch.inverse(pose).contiguous().to(cuda)      # (torch.inverse returns a column-major tensor)
focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
freqs = 2.0 ** torch.arange(pos_freq, dtype=torch.float32)
k0 = pad_in(3 * (2 * pos_freq + 1) + C)
_, _, _, feats = ops.project_gather(x, pose.to(cuda), focal, 128, 128, features=fmap, want_projection=False, pose_inv=pose_inv)
ref = encode_operand(x, freqs, k0, extra=feats)
out = torch.full((P, k0), float("nan"), device=cuda, dtype=torch.bfloat16)
fr = freqs.to(cuda)
_lib.call("nfs_g3_operand", ptr(x), ptr(pose_inv), float(focal), 128, 128, ptr(fmap), 9, 9, C, ptr(fr), pos_freq, 1, P, k0, k0, ptr(out), _stream())
torch.cuda.synchronize()
d = (out.float() - ref.float()).abs().cpu()
for p in range(P):
    bad = (d[p] > 0).nonzero().flatten().tolist()
    print(p, "bad cols", bad[:40], len(bad))
    for c in bad[:6]:
        print("    col", c, "got", float(out[p, c]), "ref", float(ref[p, c]))
print("x", x.cpu())
