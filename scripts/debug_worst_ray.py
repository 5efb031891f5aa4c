import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nerf-few-shot-limitations_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import nerf_oracle as O
from nfs_b200 import ops
torch.set_printoptions(precision=8, linewidth=220)
S, scale = 100, 100.0
g = torch.Generator().manual_seed(S * 7 + int(scale)); n = 3000
_, rd = O.lego_rays(n, seed=S)
rgb = torch.rand(n, S, 3, generator=g); den = torch.randn(n, S, 1, generator=g) * scale
z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), -1).values
gr, gd, gw = torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g) * 0.1
a, b = rgb.clone().requires_grad_(), den.clone().requires_grad_()
ref = O.render(a, b, z, rd); rg = torch.autograd.grad(list(ref), [a, b], [gr, gd, gw])
dev = "cuda"
a2, b2 = rgb.to(dev).requires_grad_(), den.to(dev).requires_grad_()
out = ops.composite(a2, b2, z.to(dev), rd.to(dev))
gg = torch.autograd.grad(list(out), [a2, b2], [gr.to(dev), gd.to(dev), gw.to(dev)])
got = gg[1].cpu()[..., 0]; r = rg[1][..., 0]
cf = O.render_backward_closed_form(rgb, den, z, rd, gr, gd, gw)[1][..., 0]
dd = lambda t: t.double()
a3, b3 = dd(rgb).requires_grad_(), dd(den).requires_grad_()
r64 = torch.autograd.grad(list(O.render(a3, b3, dd(z), dd(rd))), [b3], [dd(gr), dd(gd), dd(gw)])[0][..., 0]
floor = 0.03 * float(r.abs().max())
den_ = r.abs().max(1).values.clamp_min(floor)
err = (got - r).abs().max(1).values / den_
i = int(err.argmax()); j = int((got[i] - r[i]).abs().argmax())
print("worst ray", i, "sample", j, "err", float(err[i]), "floor", floor, "ray max", float(r[i].abs().max()))
lo, hi = max(0, j - 3), min(S, j + 4)
print("sigma ", den[i, lo:hi, 0]); print("z     ", z[i, lo:hi])
print("ref   ", r[i, lo:hi]); print("gpu   ", got[i, lo:hi]); print("cpu cf", cf[i, lo:hi]); print("fp64  ", r64[i, lo:hi])
sig = den[i, :, 0]; gaps = torch.cat([z[i, 1:] - z[i, :-1], torch.tensor([1e10])]) * rd[i].norm()
x = torch.relu(sig) * gaps; e = torch.exp(-x)
print("x     ", x[lo:hi]); print("e     ", e[lo:hi]); print("w ref ", ref[2][i, lo:hi].detach()); print("w gpu ", out[2][i, lo:hi].detach().cpu())
print("closed-form-cpu vs autograd err on this ray", float((cf[i] - r[i]).abs().max() / den_[i]))
