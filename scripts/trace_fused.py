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
    run = lambda: plan.dgrad_chain_fused(dy, save[1], P)
else:
    run = lambda: plan.run_forward_fused(x16, keep)
for _ in range(3): run()
buf = torch.zeros(2 * 18 * 1024, dtype=torch.int64, device=dev)
lib.nfs_set_debug_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.nfs_set_debug_trace(None)
lib.nfs_set_debug_flags(0)
b = buf.cpu().view(36, 1024)
ev = []
for w in range(36 if (flags & 64) else 18):          # warps 18..35 = the peer CTA (comparable only with the global timer, flag 64)
    n = int(b[w, 0])
    for i in range(n):
        v = int(b[w, 1 + i]) & ((1 << 64) - 1)
        ev.append((v >> 16, w, (v >> 12) & 15, (v >> 4) & 255, v & 15))
ev.sort()
t0 = ev[0][0]
names = {16: "mma: act_ready wait over", 17: "mma: (second trace, calibration)", 0: "mma: committed", 14: "epi: store-read wait done", 15: "epi: step begins (mask loads issued next)", 6: "mma: slab 0 issued", 7: "mma: slab 1 issued", 8: "mma: slab 2 issued", 9: "mma: slab 3 issued", 10: "mma: w_full 0", 11: "mma: w_full 1", 12: "mma: w_full 2", 13: "mma: w_full 3 / epi probe: act_ready phase complete", 1: "mma: tile ready", 2: "mma: commit acc", 3: "epi: acc_full seen", 4: "epi: math+sts done", 5: "epi: fenced"}
# second pair iteration (steady state): find the 2nd occurrence of (code 1, layer 0, tile 0)
starts = [e for e in ev if e[2] == 1 and e[3] == 0 and e[4] == 0]
lo = starts[2][0] if len(starts) > 3 else t0
hi = starts[3][0] if len(starts) > 3 else ev[-1][0]
print("cycles relative to the start of pair iteration 2; warps 2..17 = epilogue (q = w%4, cq = (w-2)//4)")
for c, w, code, l, t in ev:
    if w == 1 and code == 14: code = 16
    if w == 1 and code == 15: code = 17
    if lo <= c < hi and (w == 1 or w in (2, 17, 20, 35)) and 3 <= l <= 4:
        print("%8d  warp %2d  L%d %s  %s" % (c - lo, w, l, "AB"[t], names[code]))
print("pair iteration length:", hi - lo, "cycles")

# summary over the steady-state iteration: for every (layer, tile) when the MMA thread could start / committed, and when
# the epilogue warps of the leader CTA (and, with the global timer, of the peer CTA) saw the accumulator and had fenced
def span(lo_w, hi_w, code, l, t):
    v = [c - lo for c, w, cd, ll, tt in ev if lo <= c < hi and lo_w <= w < hi_w and cd == code and ll == l and tt == t]
    return (min(v), max(v)) if v else (-1, -1)
print("layer tile | mma ready  commit | leader: seen min..max  fenced min..max | peer: seen min..max  fenced min..max")
for l in range(9):
    for t in range(2):
        rd, cm = span(1, 2, 1, l, t), span(1, 2, 2, l, t)
        print("L%d %s | %7d %7d | %7d..%7d  %7d..%7d | %7d..%7d  %7d..%7d" % ((l, "AB"[t], rd[0], cm[0]) + span(2, 18, 3, l, t) + span(2, 18, 5, l, t)
              + span(20, 36, 3, l, t) + span(20, 36, 5, l, t)))

# per-warp phases of two steady-state tile steps (training forward / dgrad: where does the skew between warps come from?)
print("per-warp phases (cycles): lateness of 'acc_full seen' against the first warp | store-read wait | math + st.shared | fence | to next step begin")
for (l, t) in ((4, 0), (4, 1)):
    ev_lt = {}
    for c, w, code, ll, tt in ev:
        if lo <= c < hi and 2 <= w < 18 and ll == l and tt == t:
            ev_lt.setdefault(w, {})[code] = c
    nxt = {}
    for c, w, code, ll, tt in ev:
        if 2 <= w < 18 and code == 15 and ((ll == l and tt == 1 and t == 0) or (ll == l + 1 and tt == 0 and t == 1)) and lo <= c < hi + 30000:
            nxt.setdefault(w, c)
    first = min(v.get(3, 1 << 62) for v in ev_lt.values())
    print("L%d %s" % (l, "AB"[t]))
    for w in sorted(ev_lt):
        v = ev_lt[w]
        if all(k in v for k in (3, 14, 4, 5)):
            print("  warp %2d (q %d, cq %d): late %5d | %5d | %5d | %5d | %5d" % (w, w % 4, (w - 2) // 4, v[3] - first, v[14] - v[3], v[4] - v[14], v[5] - v[4],
                  nxt.get(w, v[5]) - v[5]))
