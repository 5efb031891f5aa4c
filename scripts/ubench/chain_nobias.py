"""Chain kernel with and without the bias operand (how much the bias costs)."""
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
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for label in ("with bias", "no bias"):
    print(label, "inference %.3f ms  training fwd %.3f ms" % (timed(lambda: plan.run_forward_fused(x16, False)), timed(lambda: plan.run_forward_fused(x16, True))))
    plan.b_stack = None
