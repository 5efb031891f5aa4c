"""Drop-in for /root/reference/src/models/ray_sampler.py (R7 in SURVEY.md section 8a)."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch

from nfs_b200 import ops as _ops


def get_rays(H, W, focal, c2w):
    """Pinhole rays for every pixel: rays_o, rays_d of shape (H, W, 3) on c2w's device - one kernel
    (nfs_rays_generate), bit-exact with the CPU arithmetic of ray_sampler.py:18-30 (meshgrid 'xy',
    dirs . R^T by multiply + sum).  c2w: (3,4) or (4,4) CUDA tensor."""
    rays_o, rays_d = _ops.generate_rays(H, W, focal, c2w)
    return rays_o.reshape(H, W, 3), rays_d.reshape(H, W, 3)


def sample_points_along_rays(rays_o, rays_d, near, far, N_samples, perturb=True):
    """Stratified depths and points along each ray, one fused kernel (nfs_sample_stratified).

    Accepts the reference's image-shaped rays (H, W, 3) (ray_sampler.py:47) and also flat
    (N, 3) rays, which is what the reference's own train.py:189 passes (shim (ii) of
    SURVEY.md section 8b).  Returns pts (..., S, 3), z_vals (..., S).  With perturb=True
    one torch.rand(z_vals.shape, device=...) is drawn from the global generator, as in
    ray_sampler.py:57."""
    t_rand = None
    if perturb:
        t_rand = torch.rand((*rays_o.shape[:-1], N_samples), device=rays_o.device)
    return _ops.sample_stratified(rays_o, rays_d, near, far, N_samples, t_rand=t_rand)
