"""nfs_wgrad_bf16 launch time against the number of points (fixed cost vs streaming)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for N in (256, 64):
    for P in (64, 1024, 9472, 32768, 131072, 1048576):
        u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
        v = torch.randn(P, N, device=dev).to(torch.bfloat16)
        dw = torch.zeros(N, 256, device=dev)
        cs = torch.zeros(N, device=dev)
        t0 = timed(lambda: ops.wgrad_bf16(u, v, dw, 1, 256))
        t1 = timed(lambda: ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=True))
        print("M=256 N=%3d P=%8d  %7.1f us   with colsum %7.1f us   (%.2f TB/s)" % (N, P, t0, t1, P * (256 + N) * 2 / t1 / 1e6))
