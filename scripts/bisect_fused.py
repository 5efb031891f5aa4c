import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
# needs the developer build of the library:  NFS_DEVTOOLS=1 python -m nfs_b200.build --force  (then rebuild
# without it: the production kernel compiles the bisection switches and the tracer out)
from nfs_b200 import _lib
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev)
plan = model._get_plan(); plan.refresh()
P = 4096 * 192
x16 = torch.randn(P, 64, device=dev).to(torch.bfloat16)
lib = _lib.load()
def timed(keep):
    for _ in range(3): plan.run_forward_fused(x16, keep)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10): plan.run_forward_fused(x16, keep)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for flags in (0, 1, 4, 4 + 8, 4 + 16, 4 + 32, 4 + 64, 4 + 8 + 16 + 32 + 64, 8, 16, 32, 64):
    lib.nfs_set_debug_flags(flags)
    print("flags", flags, "(1=no epilogue, 2=no weight reloads, 4=no MMAs, 8=no bias, 16=no proxy fence, 32=no st.shared, 64=no tcgen05.ld)  fwd %.3f ms   fwd+save %.3f ms" % (timed(False), timed(True)))
lib.nfs_set_debug_flags(0)
