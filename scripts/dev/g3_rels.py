import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'nerf-few-shot-limitations_b200'), os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests')]
import torch
from test_gpu_g3 import _pair, _inputs
cuda = torch.device('cuda:0')
for kwargs, P in [(dict(), 4096), (dict(pos_freq=12, dino_dim=64), 1000), (dict(dino_dim=0), 777),
                  (dict(pos_freq=6, dir_freq=2, dino_dim=16, hidden_dim=128, num_density_layers=2), 130), (dict(), 40000)]:
    ref, mod = _pair(cuda, kwargs, seed=5)
    D = kwargs.get("dino_dim", 64)
    x, d, f, t_rgb, t_den = _inputs(P, D, seed=6)
    rgb_r, den_r = ref(x, d, f)
    loss_r = ((rgb_r - t_rgb) ** 2).mean() + 0.1 * ((den_r - t_den) ** 2).mean()
    names = [k for k, _ in ref.named_parameters()]
    g_ref = torch.autograd.grad(loss_r, list(ref.parameters()))
    rgb, den = mod(x.to(cuda), d.to(cuda), f.to(cuda))
    loss = ((rgb - t_rgb.to(cuda)) ** 2).mean() + 0.1 * ((den - t_den.to(cuda)) ** 2).mean()
    grads = [t.cpu() for t in torch.autograd.grad(loss, list(mod.parameters()))]
    gmax = max(float(b.norm()) for b in g_ref)
    print(kwargs, P)
    for k, a, b in zip(names, grads, g_ref):
        if 'attention' in k or 'fusion.0' in k or 'fusion.2' in k:
            print("   %-36s rel %.4f  norm/gmax %.2e  cos %.5f" % (k, float((a - b).norm() / b.norm()), float(b.norm()) / gmax,
                  float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()))))
