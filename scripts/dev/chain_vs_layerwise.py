"""Diagnostic: fused chain vs layer-by-layer launches, per-layer difference statistics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
for P in (128, 40000):
    torch.manual_seed(2)
    mod = NeRFMLP().to(dev)
    plan = mod._get_plan(); plan.refresh()
    g = torch.Generator().manual_seed(P)
    x16 = torch.randn(P, 64, generator=g).to(torch.bfloat16).to(dev)
    x16[:, 63] = 0
    out_f, acts_f, _ = plan.run_forward_fused(x16, keep=True)
    out_f2, acts_f2, _ = plan.run_forward_fused(x16, keep=True)
    os.environ["NFS_MLP_FUSED"] = "0"
    out_l, acts_l, _ = plan.run_forward(x16, keep=True)
    os.environ["NFS_MLP_FUSED"] = "1"
    print("P", P, "deterministic:", all(torch.equal(a, b) for a, b in zip(acts_f, acts_f2)), torch.equal(out_f, out_f2))
    for i, (a, b) in enumerate(zip(acts_f, acts_l)):
        a, b = a.float(), b.float()
        d = (a - b).abs()
        k = int(d.argmax())
        r, c = divmod(k, a.shape[1])
        nrow = int((d.max(dim=1).values > 0).sum())
        print(" act %d: max|b| %.3f rms %.4f  maxdiff %.5f at (%d,%d) a=%.5f b=%.5f  frac!= %.5f rows!= %d  rel_l2 %.2e" % (
            i, float(b.abs().max()), float(b.pow(2).mean().sqrt()), float(d.max()), r, c, float(a[r, c]), float(b[r, c]),
            float((a != b).float().mean()), nrow, float((a - b).norm() / b.norm())))
    print(" out maxdiff", float((out_f - out_l).abs().max()))
