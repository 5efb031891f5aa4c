"""A/B of the chain kernels under developer flags of the -DNFS_DEVTOOLS build (NFS_B200_LIB=...libnfs_b200_devtools.so):
python scripts/dev/ab_flags.py 0 256   (256 = no look-ahead waits in the MMA issuer, 128 = no tcgen05 fence after the waits)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from nfs_b200 import _lib
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev)
plan = model._get_plan(); plan.refresh()
P = 4096 * 192
x16 = torch.randn(P, 64, device=dev).to(torch.bfloat16)
dy = torch.randn(P, 64, device=dev).to(torch.bfloat16)
lib = _lib.load()
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
_, _, save = plan.run_forward_fused(x16, True)
ref = None
for rnd in range(2):
    for f in [int(v) for v in sys.argv[1:]] or [0]:
        lib.nfs_set_debug_flags(f)
        out = plan.run_forward_fused(x16, False)[0]
        if ref is None:
            ref = out.clone()
        print("flags %4d  inference %.3f  training fwd %.3f  dgrad %.3f ms   output equal to first variant: %s" % (
            f, timed(lambda: plan.run_forward_fused(x16, False)), timed(lambda: plan.run_forward_fused(x16, True)),
            timed(lambda: plan.dgrad_chain_fused(dy, save[1], P)), bool(torch.equal(out, ref))), flush=True)
lib.nfs_set_debug_flags(0)
