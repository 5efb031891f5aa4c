"""Drop-in for /root/reference/src/models/nerf_mlp.py (R1, R4, R6, R10 of SURVEY.md 8a):
PositionalEncoding, DensityMLP, ColorMLP, NeRFWithDINO, VolumeRenderer, NeRFLoss."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch
import torch.nn as nn
import torch.nn.functional as F

from nfs_b200 import ops as _ops


class PositionalEncoding(nn.Module):
    """nerf_mlp.py:6-39.  `freq_bands` is a registered buffer (it appears in state_dict,
    unlike models.positional_encoding.PositionalEncoding); log-spaced bands only."""

    def __init__(self, num_freqs=10, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.include_input = include_input
        self.register_buffer('freq_bands', 2.0 ** torch.linspace(0.0, num_freqs - 1, num_freqs))

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("PositionalEncoding: gradients w.r.t. coordinates are not supported")
        return _ops.posenc(x, self.freq_bands, self.include_input)

    def get_output_dim(self, input_dim):
        return input_dim * (2 * self.num_freqs) + (input_dim if self.include_input else 0)


class VolumeRenderer(nn.Module):
    """Alpha compositing along rays, nerf_mlp.py:160-215, as one CUDA kernel forward
    (nfs_composite_fwd) and one backward (nfs_composite_bwd, recomputing alpha / T)."""

    def __init__(self):
        super().__init__()

    def forward(self, rgb, density, z_vals, rays_d, noise_std=0.0, white_bkgd=False):
        """rgb (N,S,3), density (N,S,1), z_vals (N,S), rays_d (N,3) ->
        rgb_rendered (N,3), depth_rendered (N,), weights (N,S).
        Noise is drawn (one randn_like(density) from the global generator of the input's
        device) only when noise_std > 0 and the module is in training mode (:188-190)."""
        noise = None
        if noise_std > 0.0 and self.training:
            noise = torch.randn_like(density)
        return _ops.composite(rgb, density, z_vals, rays_d, noise=noise, noise_std=noise_std,
                              white_bkgd=white_bkgd)


class NeRFLoss(nn.Module):
    """rgb MSE [+ depth L1 when the target carries depth] + mean(weights^2), nerf_mlp.py:217-258.
    A handful of scalar reductions; which terms are present decides which upstream gradients
    (g_rgb, g_depth, g_weights) reach the compositing backward."""

    def __init__(self, rgb_weight=1.0, depth_weight=0.1, regularization_weight=0.01):
        super().__init__()
        self.rgb_weight = rgb_weight
        self.depth_weight = depth_weight
        self.reg_weight = regularization_weight

    def forward(self, predictions, targets, weights=None):
        losses = {'rgb': F.mse_loss(predictions['rgb'], targets['rgb'])}
        if 'depth' in targets:
            losses['depth'] = F.l1_loss(predictions['depth'], targets['depth'])
        if 'weights' in predictions:
            losses['regularization'] = torch.mean(predictions['weights'] ** 2)
        total = self.rgb_weight * losses['rgb']
        if 'depth' in losses:
            total = total + self.depth_weight * losses['depth']
        if 'regularization' in losses:
            total = total + self.reg_weight * losses['regularization']
        losses['total'] = total
        return losses
