"""Timeline of CTA 0 of the fused MLP kernel (developer tool): prints, for the first pair of tiles,
when the MMA thread saw each tile ready / committed its accumulator and when the epilogue warps
woke, finished their math and signalled the next layer."""
import os, sys
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
keep = len(sys.argv) > 1 and sys.argv[1] == "save"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib.nfs_set_debug_flags(flags)
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
if bwd:
    out, acts, save = plan.run_forward_fused(x16, True)
    dy = torch.randn(P, 64, device=dev).to(torch.bfloat16)
    run = lambda: plan.dgrad_chain_fused(dy, save, P)
else:
    run = lambda: plan.run_forward_fused(x16, keep)
for _ in range(3): run()
buf = torch.zeros(18 * 1024, dtype=torch.int64, device=dev)
lib.nfs_set_debug_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.nfs_set_debug_trace(None)
lib.nfs_set_debug_flags(0)
b = buf.cpu().view(18, 1024)
ev = []
for w in range(18):
    n = int(b[w, 0])
    for i in range(n):
        v = int(b[w, 1 + i]) & ((1 << 64) - 1)
        ev.append((v >> 16, w, (v >> 12) & 15, (v >> 4) & 255, v & 15))
ev.sort()
t0 = ev[0][0]
names = {14: "epi: store-read wait done", 15: "epi: step begins (mask loads issued next)", 6: "mma: slab 0 issued", 7: "mma: slab 1 issued", 8: "mma: slab 2 issued", 9: "mma: slab 3 issued", 10: "mma: w_full 0", 11: "mma: w_full 1", 12: "mma: w_full 2", 13: "mma: w_full 3", 1: "mma: tile ready", 2: "mma: commit acc", 3: "epi: acc_full seen", 4: "epi: math+sts done", 5: "epi: fenced"}
# second pair iteration (steady state): find the 2nd occurrence of (code 1, layer 0, tile 0)
starts = [e for e in ev if e[2] == 1 and e[3] == 0 and e[4] == 0]
lo = starts[2][0] if len(starts) > 3 else t0
hi = starts[3][0] if len(starts) > 3 else ev[-1][0]
print("cycles relative to the start of pair iteration 2; warps 2..17 = epilogue (q = w%4, cq = (w-2)//4)")
for c, w, code, l, t in ev:
    if lo <= c < hi and (w == 1 or w in (2,)) and 3 <= l <= 4:
        print("%8d  warp %2d  L%d %s  %s" % (c - lo, w, l, "AB"[t], names[code]))
print("pair iteration length:", hi - lo, "cycles")
