"""Dev: where does a mixed-width chain differ from torch?  Prints, per saved layer, the 128-row tiles whose error is
above tolerance.  usage: chain_mixed.py [P] [dims like 192,256,256,128,64]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import _lib
from nfs_b200._lib import ptr
from nfs_b200.mlp import bias_terms
from nfs_b200.ops import _stream

P = int(sys.argv[1]) if len(sys.argv) > 1 else 40001
w_list = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "192,256,256,128,64").split(",")]
dims = list(zip(w_list[:-1], w_list[1:]))
cuda = torch.device("cuda:0")
g = torch.Generator().manual_seed(P)
Ws = [(torch.randn(n, k, generator=g) * (2.0 / k) ** 0.5).to(torch.bfloat16) for k, n in dims]
bs = [torch.randn(n, generator=g) * 0.1 for k, n in dims]
x = torch.randn(P, dims[0][0], generator=g).to(torch.bfloat16)
i32 = lambda v: (ctypes.c_int32 * len(v))(*v)
rows = (P + 127) // 128 * 128
w = torch.zeros(sum(m.shape[0] for m in Ws), 256, dtype=torch.bfloat16)
r, row0 = 0, []
for m in Ws:
    w[r:r + m.shape[0], :m.shape[1]] = m
    row0.append(r); r += m.shape[0]
w = w.to(cuda)
bt = bias_terms(torch.cat(bs).to(cuda))
L = len(dims)
wid = max(n for k, n in dims[:-1])
save = torch.full((L - 1, rows, wid), float("nan"), device=cuda, dtype=torch.bfloat16)
bits = torch.zeros((L - 1, rows, 8), device=cuda, dtype=torch.int32)
out = torch.empty((P, 2), device=cuda, dtype=torch.float32)
xc = x.to(cuda)
_lib.call("nfs_mlp_chain", ptr(xc), P, L, i32([k for k, n in dims]), i32([n for k, n in dims]), i32([1] * (L - 1) + [0]),
          i32(row0), ptr(w), w.shape[0], ptr(bt), None, 0, None, ptr(save), ptr(bits), rows, ptr(out), 2, _stream())
torch.cuda.synchronize()
h = x.float()
for l in range(L - 1):
    h = torch.relu(h @ Ws[l].float().T + bs[l]).to(torch.bfloat16).float()
    n = dims[l][1]
    got = save[l, :P, :n].float().cpu()
    err = (got - h).abs()
    err[torch.isnan(err)] = 1e9
    tile_err = torch.stack([err[t * 128:(t + 1) * 128].max() for t in range((P + 127) // 128)])
    bad = (tile_err > 0.02 * h.abs().max() + 1e-3).nonzero().flatten().tolist()
    print("layer", l, "width", n, "max err", float(err.max()), "ref max", float(h.abs().max()), "bad tiles", bad[:40], len(bad))
    if bad:
        t = bad[0]
        e = err[t * 128:(t + 1) * 128]
        rr, cc = (e > 0.02 * h.abs().max() + 1e-3).nonzero(as_tuple=True)
        print("   first bad tile", t, "rows", sorted(set(rr.tolist()))[:10], "cols", sorted(set(cc.tolist()))[:20], len(rr))
    h = got.nan_to_num(0.0)
ref = (h @ Ws[-1].float().T + bs[-1])[:, :2]
print("head err", float((out.cpu() - ref).abs().max()))
