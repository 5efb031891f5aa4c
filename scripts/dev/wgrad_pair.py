"""CTA-pair weight-gradient body (wgrad_pair_body.cuh, NFS_WGRAD_PAIR=1) against the single-CTA kernel: results on a few
shapes, then throughput per SM at full and reduced grids."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(u, v, pair, colsum_of_v, n_valid=256):
    dw = torch.zeros(256, 256, device=dev)
    cs = torch.zeros(256, device=dev)
    if pair:
        os.environ["NFS_WGRAD_PAIR"] = "1"
    else:
        os.environ.pop("NFS_WGRAD_PAIR", None)
    ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=colsum_of_v, n_valid=n_valid)
    torch.cuda.synchronize()
    return dw, cs


for P in (64, 1000, 4096, 70001, 786432):
    for csv in (True, False):
        for nv in (256, 200):
            u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
            v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
            d0, c0 = run(u, v, False, csv, nv)
            d1, c1 = run(u, v, True, csv, nv)
            ref = (v.float().t() @ u.float())
            ref[nv:] = 0
            cref = (v if csv else u).float().sum(0)
            if csv:
                cref[nv:] = 0
            print("P=%7d colsum_of_v=%d n_valid=%d  dw: pair-vs-single %.2e  pair-vs-fp32 %.2e  single-vs-fp32 %.2e | colsum: %.2e %.2e %.2e" % (
                P, csv, nv, float((d1 - d0).norm() / d0.norm()), float((d1 - ref).norm() / ref.norm()), float((d0 - ref).norm() / ref.norm()),
                float((c1 - c0).norm() / c0.norm()), float((c1 - cref).norm() / cref.norm()), float((c0 - cref).norm() / cref.norm())), flush=True)

for P in (64, 1000, 70001, 786432):
    for csv in (True, False):
        u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
        v = torch.randn(P, 64, device=dev).to(torch.bfloat16)
        res = []
        for pair in (False, True):
            dw = torch.zeros(64, 256, device=dev)
            cs = torch.zeros(256, device=dev)
            if pair:
                os.environ["NFS_WGRAD_PAIR"] = "1"
            else:
                os.environ.pop("NFS_WGRAD_PAIR", None)
            ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=csv, n_valid=63)
            torch.cuda.synchronize()
            res.append((dw, cs))
        ref = (v.float().t() @ u.float()); ref[63:] = 0
        cref = torch.zeros(256, device=dev)
        if csv:
            cref[:63] = v.float().sum(0)[:63]
        else:
            cref = u.float().sum(0)
        (d0, c0), (d1, c1) = res
        print("N=64 P=%7d colsum_of_v=%d  dw: pair-vs-single %.2e  pair-vs-fp32 %.2e | colsum: pair-vs-single %.2e pair-vs-fp32 %.2e" % (
            P, csv, float((d1 - d0).norm() / d0.norm()), float((d1 - ref).norm() / ref.norm()),
            float((c1 - c0).norm() / c0.norm()), float((c1 - cref).norm() / cref.norm())), flush=True)

P = 786432
for pair in (0, 1):
    for csv in (True, False):
        for grid in (148, 36, 16):
            u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
            v = torch.randn(P, 64, device=dev).to(torch.bfloat16)
            dw = torch.zeros(64, 256, device=dev)
            cs = torch.zeros(256, device=dev)
            os.environ["NFS_WGRAD_GRID"] = str(grid)
            if pair:
                os.environ["NFS_WGRAD_PAIR"] = "1"
            else:
                os.environ.pop("NFS_WGRAD_PAIR", None)
            f = lambda: ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=csv, n_valid=63)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            for _ in range(10):
                f()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print("N=64 pair=%d colsum_of_v=%d grid=%3d %.3f ms  %.1f ns per slab and CTA" % (pair, csv, grid, ms, ms * 1e6 / (P / 64) * grid), flush=True)
os.environ.pop("NFS_WGRAD_GRID", None)
if len(sys.argv) < 2:
    sys.exit(0)
u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
dw = torch.zeros(256, 256, device=dev)
cs = torch.zeros(256, device=dev)
for pair in (0, 1):
    for dbg in (0, 2, 1, 3):
        for grid in (148, 74, 36, 16):
            for colsum in (cs, None):
                os.environ["NFS_WGRAD_GRID"] = str(grid)
                os.environ["NFS_WGRAD_DBG"] = str(dbg)
                if pair:
                    os.environ["NFS_WGRAD_PAIR"] = "1"
                else:
                    os.environ.pop("NFS_WGRAD_PAIR", None)
                f = lambda: ops.wgrad_bf16(u, v, dw, 1, 256, colsum=colsum, colsum_of_v=True)
                for _ in range(3):
                    f()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record()
                for _ in range(10):
                    f()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 10
                gb = P * 1024 / 1e9
                print("pair=%d dbg=%d grid=%3d colsum=%-5s %.3f ms  %.0f GB/s total  %.1f GB/s per SM" % (
                    pair, dbg, grid, colsum is not None, ms, gb / ms * 1e3, gb / ms * 1e3 / grid), flush=True)
