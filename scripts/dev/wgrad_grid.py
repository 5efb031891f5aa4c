"""How does nfs_wgrad_bf16 scale when only a subset of the SMs runs it?  (Design input for the merged backward
kernel: dgrad chain on some CTA pairs, weight-gradient consumers on the rest.)  Times P = 786 432 points, M = N = 256
at grids 148 ... 16 (NFS_WGRAD_GRID), operands from HBM (cold) and a 32 768-point launch repeated so that its
operands (33 MB) stay in L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")


def run(P, grid, reps):
    u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
    v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
    dw = torch.zeros(256, 256, device=dev)
    cs = torch.zeros(256, device=dev)
    os.environ["NFS_WGRAD_GRID"] = str(grid)
    for _ in range(3):
        ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(reps):
        ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gb = P * 1024 / 1e9
    print("P=%8d grid=%3d  %.3f ms  %.0f GB/s total  %.1f GB/s per SM" % (P, grid, ms, gb / ms * 1e3, gb / ms * 1e3 / grid))


for grid in (148, 111, 74, 56, 37, 16):
    run(786432, grid, 10)
for grid in (148, 74, 37, 16):
    run(32768, grid, 50)
