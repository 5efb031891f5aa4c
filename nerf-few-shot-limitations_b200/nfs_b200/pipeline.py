"""The render path end to end: stratified sampling -> (fused encoding +) MLP -> compositing
[-> inverse-CDF resampling -> MLP -> compositing], i.e. what NeRFDINOTrainer.render_rays
(/root/reference/src/training/train.py:188-242) does for the baseline model, with the
hierarchical pass of utils.ray_utils.hierarchical_sampling (ray_utils.py:86-143) wired in the
way BASELINE.json config 3 / 5 describe (64 coarse + 128 importance samples = 192 fine).

Every stage is one of the library's CUDA kernels; tensors between stages stay in HBM in the
layout the next kernel wants (the MLP emits packed [r,g,b,sigma] rows, which is exactly the
packed operand of the compositing kernel, so nothing is sliced or copied in between).
"""
import torch

from . import ops


def render_rays(model, freq_bands, rays_o, rays_d, near, far, n_coarse=64, n_importance=0, perturb=True,
                white_bkgd=False, t_rand=None, u=None):
    """model: models.nerf_model.NeRFMLP (or any module with forward_points(points, freq_bands)
    -> (...,4) [rgb|sigma]).  rays (N,3).  Returns a dict with rgb/depth/weights/z of the last
    pass ('rgb', ...) and of the coarse pass ('rgb_coarse', ...) when n_importance > 0.
    t_rand (N,n_coarse) / u (N,n_importance) override the draws (parity tests); otherwise they come
    from the global CUDA generator in the order the reference draws them."""
    N = rays_o.shape[0]
    dev = rays_o.device
    if perturb and t_rand is None:
        t_rand = torch.rand(N, n_coarse, device=dev)
    pts, z = ops.sample_stratified(rays_o, rays_d, near, far, n_coarse, t_rand=t_rand if perturb else None)
    raw = model.forward_points(pts.reshape(-1, 3), freq_bands).reshape(N, n_coarse, 4)
    rgb, depth, weights = ops.composite_packed(raw, z, rays_d, white_bkgd=white_bkgd, want_aux=True)
    out = {"rgb": rgb, "depth": depth, "weights": weights, "z_vals": z}
    if n_importance > 0:
        out.update(rgb_coarse=rgb, depth_coarse=depth, weights_coarse=weights, z_coarse=z)
        with torch.no_grad():
            w_in = weights.detach()[:, :-1].contiguous()      # M = S-1 bins between the S coarse depths
            if u is None:
                u = torch.rand(N, n_importance, device=dev) if perturb else \
                    torch.linspace(0., 1., n_importance).to(dev)
            pts_f, z_f = ops.sample_hierarchical(rays_o, rays_d, z, w_in, n_importance, u=u)
        S = n_coarse + n_importance
        raw_f = model.forward_points(pts_f.reshape(-1, 3), freq_bands).reshape(N, S, 4)
        rgb_f, depth_f, w_f = ops.composite_packed(raw_f, z_f, rays_d, white_bkgd=white_bkgd, want_aux=True)
        out.update(rgb=rgb_f, depth=depth_f, weights=w_f, z_vals=z_f)
    return out


def train_step(model, optimizer, freq_bands, rays_o, rays_d, target, near, far, n_coarse=64, n_importance=128,
               perturb=True, loss_scale=1.0, allreduce=None):
    """One optimisation step of BASELINE config 3: render (coarse + fine), MSE on both passes
    (train.py:36-44 uses rgb MSE only), backward through the compositing and MLP kernels,
    optional gradient all-reduce, fused Adam.  Returns the (detached) loss tensor - no host sync."""
    optimizer.zero_grad()
    out = render_rays(model, freq_bands, rays_o, rays_d, near, far, n_coarse, n_importance, perturb)
    loss = torch.mean((out["rgb"] - target) ** 2)
    if n_importance > 0:
        loss = loss + torch.mean((out["rgb_coarse"] - target) ** 2)
    (loss * loss_scale if loss_scale != 1.0 else loss).backward()
    if hasattr(optimizer, "gather_grads"):
        g = optimizer.gather_grads()
        scale = 1.0
        if allreduce is not None:
            allreduce(g)
        optimizer.step(grad_scale=scale, gathered=True)
    else:
        optimizer.step()
    return loss.detach()


@torch.no_grad()
def render_image(model, freq_bands, rays_o, rays_d, near, far, n_coarse=64, n_importance=128, chunk=65536,
                 white_bkgd=False):
    """Full-frame render (BASELINE config 5): rays (R,3) in chunks, eval mode (perturb=False)."""
    outs = []
    for i in range(0, rays_o.shape[0], chunk):
        o = render_rays(model, freq_bands, rays_o[i:i + chunk], rays_d[i:i + chunk], near, far, n_coarse,
                        n_importance, perturb=False, white_bkgd=white_bkgd)
        outs.append(o["rgb"])
    return torch.cat(outs, 0) if outs else rays_o.new_zeros((0, 3))
