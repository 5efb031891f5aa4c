"""A/B of the quotient in K1's backward: __fdividef (default build) against __fdiv_rn (python -m nfs_b200.build
--variant exactdiv).  Run once per library:  NFS_B200_LIB=<path> python scripts/dev/ab_k1_div.py
Prints the time of the staged backward at the benchmark size and its error against the CPU oracle (fp32 autograd) and
against the fp64 closed form, on 131 072 rays of the benchmark distribution and on the sigma ~ 100 N(0,1) stress case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from helpers import per_ray_err
from nfs_b200 import _lib, ops
from oracle import nerf_oracle as O

dev = torch.device("cuda:0")
print("library:", _lib.LIB_PATH)
torch.manual_seed(0)
for scale in (10.0, 100.0):
    n, S = 1 << 17, 64
    d = bench.make_inputs(n, S, seed=1, device=None)
    den = d["density"] * (scale / 10.0)
    g_rgb, g_depth = torch.randn(n, 3) / n, torch.randn(n) / n
    a, b = d["rgb"].clone().requires_grad_(), den.clone().requires_grad_()
    ref = O.render(a, b, d["z"], d["rays_d"])
    ref_g = torch.autograd.grad([ref[0], ref[1]], [a, b], [g_rgb, g_depth])
    f64 = O.render_backward_closed_form(d["rgb"].double(), den.double(), d["z"].double(), d["rays_d"].double(), g_rgb.double(),
                                        g_depth.double())
    a2, b2 = d["rgb"].to(dev).requires_grad_(), den.to(dev).requires_grad_()
    out = ops.composite(a2, b2, d["z"].to(dev), d["rays_d"].to(dev))
    got = torch.autograd.grad([out[0], out[1]], [a2, b2], [g_rgb.to(dev), g_depth.to(dev)])
    print("sigma scale %5.0f: d_density per-ray err vs fp32 autograd %.3e (floor 3%%: %.3e) | vs fp64 closed form %.3e (floor 3%%: %.3e)"
          " | reference fp32 autograd vs fp64: %.3e (floor 3%%: %.3e)" % (
              scale, per_ray_err(got[1], ref_g[1]), per_ray_err(got[1], ref_g[1], floor_frac=0.03),
              per_ray_err(got[1], f64[1].float()), per_ray_err(got[1], f64[1].float(), floor_frac=0.03),
              per_ray_err(ref_g[1], f64[1].float()), per_ray_err(ref_g[1], f64[1].float(), floor_frac=0.03)))
n, S = 1 << 20, 64
d = bench.make_inputs(n, S, seed=0, device=dev)
rgb, den = d["rgb"].requires_grad_(), d["density"].requires_grad_()
out = ops.composite(rgb, den, d["z"], d["rays_d"])
g_rgb, g_depth = torch.randn(n, 3, device=dev) / n, torch.randn(n, device=dev) / n
for _ in range(5):
    torch.autograd.grad([out[0], out[1]], [rgb, den], [g_rgb, g_depth], retain_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    torch.autograd.grad([out[0], out[1]], [rgb, den], [g_rgb, g_depth], retain_graph=True)
e1.record()
torch.cuda.synchronize()
print("staged backward, 2^20 rays x 64: %.4f ms" % (e0.elapsed_time(e1) / 20))
