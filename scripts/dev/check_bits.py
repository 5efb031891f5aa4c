"""Compare the ReLU mask bits written by the training forward with (saved activation > 0), element by element."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
torch.manual_seed(2)
mod = NeRFMLP().to(dev)
plan = mod._get_plan(); plan.refresh()
for P in (257, 40000, 113669):
    g = torch.Generator().manual_seed(P)
    x16 = torch.randn(P, 64, generator=g).to(torch.bfloat16).to(dev)
    x16[:, 63] = 0
    out, acts, (save, bits) = plan.run_forward_fused(x16, keep=True)
    torch.cuda.synchronize()
    L, rows, _ = bits.shape
    w = bits[:, :P].to(torch.int64) & 0xFFFFFFFF                      # [L, P, 8]
    j = torch.arange(16, device=dev)
    even = (w[..., None] >> (15 - j)) & 1                             # [L,P,8,16] column 32w + 2j
    odd = (w[..., None] >> (31 - j)) & 1
    mask = torch.stack([even, odd], -1).reshape(L, P, 256).bool()
    ref = save[:, :P] > 0
    bad = mask != ref
    print("P=%d: %d mismatches of %d" % (P, int(bad.sum()), bad.numel()))
    if bad.any():
        idx = bad.nonzero()[:10]
        for l, p, c in idx.tolist():
            print("   layer %d row %d col %d: bit %d, activation %r" % (l, p, c, int(mask[l, p, c]), float(save[l, p, c])))
        print("   per layer:", bad.sum((1, 2)).tolist(), " rows with mismatches (first 10):", bad.any(2).any(0).nonzero()[:10].flatten().tolist())
