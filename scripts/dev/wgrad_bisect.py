"""Bisection of the weight-gradient kernel's per-SM throughput (NFS_WGRAD_DBG: 1 = no MMAs, 2 = no column sums) at
reduced grids: what bounds a CTA at ~56 GB/s when HBM is not the limit?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")
P = 786432
u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
dw = torch.zeros(256, 256, device=dev)
cs = torch.zeros(256, device=dev)
for dbg in (0, 2, 1, 3):
    for grid in (148, 74, 37):
        for colsum in (cs, None):
            os.environ["NFS_WGRAD_GRID"] = str(grid)
            os.environ["NFS_WGRAD_DBG"] = str(dbg)
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
            print("dbg=%d grid=%3d colsum=%-5s %.3f ms  %.0f GB/s total  %.1f GB/s per SM" % (
                dbg, grid, colsum is not None, ms, gb / ms * 1e3, gb / ms * 1e3 / grid), flush=True)
