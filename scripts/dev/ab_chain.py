"""A/B timing of the chain kernel (inference / training forward / dgrad) under environment switches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from models.nerf_model import NeRFMLP
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev)
plan = model._get_plan(); plan.refresh()
P = 4096 * 192
x16 = torch.randn(P, 64, device=dev).to(torch.bfloat16)
dy = torch.randn(P, 64, device=dev).to(torch.bfloat16)
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
_, _, save = plan.run_forward_fused(x16, True)
for rnd in range(2):
    for v in sys.argv[1:] or ["-"]:
        env = dict(kv.split("=") for kv in v.split(",")) if v != "-" else {}
        os.environ.update(env)
        print("%-24s inference %.3f  training fwd %.3f  dgrad %.3f ms" % (v, timed(lambda: plan.run_forward_fused(x16, False)),
              timed(lambda: plan.run_forward_fused(x16, True)), timed(lambda: plan.dgrad_chain_fused(dy, save[1], P))))
        for k in env: del os.environ[k]
