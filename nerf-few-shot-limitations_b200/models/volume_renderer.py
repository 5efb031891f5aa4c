"""Drop-in for /root/reference/src/models/volume_renderer.py (R2 in SURVEY.md section 8a)."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:  # imported by bare name (src/models on sys.path)
    import _bootstrap  # noqa: F401

import torch

from nfs_b200 import ops as _ops


def volume_render_radiance(rgb_sigma, z_vals, rays_d, noise_std=0.0):
    """Alpha-composite packed [R,G,B,sigma] samples into an image.

    Same contract as the reference function (volume_renderer.py:4-43):
      rgb_sigma (..., S, 4), z_vals (..., S), rays_d (..., 3) -> rgb_map (..., 3).
    When noise_std > 0 the reference perturbs sigma IN PLACE on a view of its input
    (`sigma += noise_std * randn_like(sigma)`, :28-29) regardless of train/eval; that
    side effect and the RNG draw (one randn of sigma's shape from the global generator
    of the input's device) are kept.  Everything else is one CUDA kernel
    (nfs_composite_fwd / nfs_composite_bwd, packed layout).
    """
    if noise_std > 0.0:
        sigma = rgb_sigma[..., 3]
        sigma += noise_std * torch.randn_like(sigma)
    return _ops.composite_packed(rgb_sigma, z_vals, rays_d)
