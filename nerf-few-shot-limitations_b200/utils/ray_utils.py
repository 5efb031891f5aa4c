"""Drop-in for /root/reference/src/utils/ray_utils.py (R8, R9 and the 8f helpers)."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch

from nfs_b200 import ops as _ops


def get_rays(H, W, focal, pose):
    """(H, W, 3) ray origins / directions for a (4,4) camera-to-world pose (ray_utils.py:4-37) - one kernel
    (nfs_rays_generate), bit-exact with the reference's CPU arithmetic.  The reference builds the pixel grid
    on the CPU (ray_utils.py:18-22), which breaks for a pose on any other device; here
    the rays are produced on pose.device, which must be a CUDA device."""
    rays_o, rays_d = _ops.generate_rays(H, W, focal, pose)
    return rays_o.reshape(H, W, 3), rays_d.reshape(H, W, 3)


def sample_points_along_rays(rays_o, rays_d, near, far, N_samples, perturb=True, lindisp=False):
    """(N,3) rays -> pts (N,S,3), z_vals (N,S); ray_utils.py:39-84 as one kernel."""
    t_rand = None
    if perturb:
        t_rand = torch.rand((*rays_o.shape[:-1], N_samples), device=rays_o.device)
    return _ops.sample_stratified(rays_o, rays_d, near, far, N_samples, t_rand=t_rand, lindisp=lindisp)


def hierarchical_sampling(rays_o, rays_d, z_vals, weights, N_importance, perturb=True):
    """Inverse-CDF importance sampling (ray_utils.py:86-143) as one warp-per-ray kernel.

    Like the reference this only works when weights has one entry fewer than z_vals
    (z_vals are used as the M+1 bin edges of M weights; the documented (N,S)/(N,S) call
    raises in the reference too - SURVEY.md section 0.4).  Returns
    pts (N, S+N_importance, 3) and the sorted z_vals (N, S+N_importance)."""
    N_rays = z_vals.shape[0]
    device = z_vals.device
    if perturb:
        u = torch.rand(N_rays, N_importance, device=device)
    else:
        u = torch.linspace(0., 1., N_importance).to(device)   # CPU linspace: see ops.stratified_tables
    return _ops.sample_hierarchical(rays_o, rays_d, z_vals, weights, N_importance, u=u)


def get_ray_batch(rays_o, rays_d, batch_size=1024):
    """Yield (rays_o, rays_d, pixel_indices) slices of an (H,W,3) ray image
    (ray_utils.py:145-174); the index tensor follows the rays' device."""
    H, W = rays_o.shape[:2]
    N_rays = H * W
    rays_o_flat = rays_o.reshape(-1, 3)
    rays_d_flat = rays_d.reshape(-1, 3)
    indices = torch.arange(N_rays, device=rays_o.device)
    for i in range(0, N_rays, batch_size):
        end_i = min(i + batch_size, N_rays)
        yield rays_o_flat[i:end_i], rays_d_flat[i:end_i], indices[i:end_i]


def project_points_to_image(points_3d, pose, focal, H, W):
    """World points -> normalised [-1,1] image coordinates, camera depth, in-front mask
    (ray_utils.py:176-210) as one kernel (nfs_project_gather); nfs_b200.ops.project_gather also does the
    bilinear feature lookup that follows it in the callers (train.py:212-221) in the same pass."""
    return _ops.project_gather(points_3d, pose, focal, H, W)
