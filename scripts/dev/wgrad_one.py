import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import ops
dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
u = torch.randn(P, 256, device=dev).to(torch.bfloat16)
v = torch.randn(P, 256, device=dev).to(torch.bfloat16)
dw = torch.zeros(256, 256, device=dev)
cs = torch.zeros(256, device=dev)
for _ in range(4):
    ops.wgrad_bf16(u, v, dw, 1, 256, colsum=cs, colsum_of_v=True)
torch.cuda.synchronize()
