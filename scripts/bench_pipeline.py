"""Developer benchmark: BASELINE config 3 (train step, 4096 rays, 64 + 192 evaluations/ray) and
config 5 (full-frame render, 192 samples/ray) with a per-kernel breakdown from CUDA events.
Prints one JSON object; bench.py remains the contract benchmark."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200")]
import torch
from oracle import nerf_oracle as O
from nfs_b200 import pipeline, ops, mlp
from nfs_b200.optim import FusedAdam
from models.nerf_model import NeRFMLP

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeRFMLP().to(dev).train()
opt = FusedAdam(model.parameters(), lr=5e-4)
bands = O.frequency_bands(10)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ro, rd = O.lego_rays(N, seed=0)
ro, rd = ro.to(dev), rd.to(dev)
target = torch.rand(N, 3, device=dev)


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (time.perf_counter() - t0) * 1e3 / reps


res = {"rays": N}
ms, wall = timed(lambda: pipeline.train_step(model, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128))
res["train_step_ms"] = ms; res["train_step_wall_ms"] = wall; res["train_rays_per_s"] = N / ms * 1e3
flop = N * 256 * 2823168
res["train_tflops"] = flop / ms / 1e9
# pieces
P = N * 192
x = torch.randn(P, 256, device=dev).to(torch.bfloat16)
plan = model._get_plan(); plan.refresh()
w = plan.packed[1]
res["linear_256x256_ms"] = timed(lambda: ops.linear_bf16(x, w.w16, w.bias, act=1))[0]
res["linear_256x256_GBs"] = P * 256 * 2 * 2 / res["linear_256x256_ms"] / 1e6
res["linear_256x256_tflops"] = 2 * P * 256 * 256 / res["linear_256x256_ms"] / 1e9
res["dgrad_256x256_masked_ms"] = timed(lambda: ops.linear_bf16(x, w.w16t, None, act=0, relu_mask_src=x))[0]
dw = torch.zeros(256, 256, device=dev); db = torch.zeros(256, device=dev)
res["wgrad_256x256_ms"] = timed(lambda: ops.wgrad_bf16(x, x, dw, 1, 256, colsum=db))[0]
res["wgrad_256x256_GBs"] = P * 256 * 2 * 2 / res["wgrad_256x256_ms"] / 1e6
pts = torch.randn(P, 3, device=dev)
res["posenc_bf16_ms"] = timed(lambda: mlp.encode_operand(pts, bands, 64))[0]
with torch.no_grad():
    res["mlp_fwd_nograd_ms"] = timed(lambda: model.forward_points(pts, bands))[0]
    res["mlp_fwd_tflops"] = P * 951808 / res["mlp_fwd_nograd_ms"] / 1e9
out = model.forward_points(pts, bands)
gout = torch.randn_like(out)
def fb():
    o = model.forward_points(pts, bands)
    torch.autograd.grad(o, list(model.parameters()), gout)
res["mlp_fwd_bwd_ms"] = timed(fb, reps=10)[0]
res["mlp_fwd_bwd_tflops"] = P * 2823168 / res["mlp_fwd_bwd_ms"] / 1e9
# full-frame render, config 5 (one GPU's share when sharded is R/G rays)
R = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
ro2, rd2 = O.lego_rays(R, seed=1)
ro2, rd2 = ro2.to(dev), rd2.to(dev)
model.eval()
ms, wall = timed(lambda: pipeline.render_image(model, bands, ro2, rd2, 2.0, 6.0, 64, 128, chunk=65536), reps=3, warm=1)
res["render_rays"] = R; res["render_ms"] = ms; res["render_rays_per_s"] = R / ms * 1e3
res["render_tflops"] = R * 256 * 951808 / ms / 1e9
print(json.dumps(res, indent=1))
