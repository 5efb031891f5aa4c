import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")
tr = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["NFS_WG_TRACE"] = str(tr.data_ptr())
for P in (64, 32768, 1048576):
    u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
    v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
    dw = torch.zeros(256, 256, device=dev)
    cs = torch.zeros(256, device=dev)
    for _ in range(3):
        tr.zero_()
        ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=True)
        torch.cuda.synchronize()
    t = tr.cpu().tolist()
    names = ["start", "setup done", "first slab landed", "last MMA committed", "drain: enter", "drain: acc ready", "r0 staged", "r0 smem read", "", "r1 staged", "r1 smem read", "", "", "", "bulk complete", "dealloc"]
    print("P =", P)
    for i, (n, x) in enumerate(zip(names, t)):
        if x:
            print("   %-20s %8d cycles" % (n, x - t[0]))
