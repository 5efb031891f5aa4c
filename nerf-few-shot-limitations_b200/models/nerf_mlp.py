"""Drop-in for /root/reference/src/models/nerf_mlp.py (R1, R4, R6, R10 of SURVEY.md 8a):
PositionalEncoding, DensityMLP, ColorMLP, NeRFWithDINO, VolumeRenderer, NeRFLoss."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch
import torch.nn as nn
import torch.nn.functional as F

from nfs_b200 import mlp_g3 as _g3
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


class DensityMLP(nn.Module):
    """num_layers x (Linear + ReLU) -> relu(density_head), feature_head; nerf_mlp.py:41-66.
    Same sub-module names (`density_layers.{0,2,..}`, `density_head`, `feature_head`)."""

    def __init__(self, input_dim, hidden_dim=256, num_layers=4):
        super().__init__()
        layers = []
        for i in range(num_layers):
            layers.append(nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim))
            layers.append(nn.ReLU(inplace=True))
        self.density_layers = nn.Sequential(*layers)
        self.density_head = nn.Linear(hidden_dim, 1)
        self.feature_head = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, x):
        h = x
        for layer in self.density_layers:
            if isinstance(layer, nn.Linear):
                h = _g3.dense(self, layer, h, "relu")
        return _g3.dense(self, self.density_head, h, "relu"), _g3.dense(self, self.feature_head, h)


class ColorMLP(nn.Module):
    """[features | view_dirs] -> hidden -> hidden/2 -> sigmoid rgb; nerf_mlp.py:68-84."""

    def __init__(self, feature_dim, dir_dim, hidden_dim=128):
        super().__init__()
        self.color_layers = nn.Sequential(
            nn.Linear(feature_dim + dir_dim, hidden_dim), nn.ReLU(inplace=True),
            nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(inplace=True),
            nn.Linear(hidden_dim // 2, 3), nn.Sigmoid())

    def forward(self, features, view_dirs):
        h = torch.cat([features, view_dirs], dim=-1)
        h = _g3.dense(self, self.color_layers[0], h, "relu")
        h = _g3.dense(self, self.color_layers[2], h, "relu")
        return _g3.dense(self, self.color_layers[4], h, "sigmoid")


class NeRFWithDINO(nn.Module):
    """nerf_mlp.py:86-158: positional encodings -> NeRFDINOFusion -> DensityMLP -> ColorMLP.
    Constructor arguments, sub-module names and parameter shapes are the reference's (818 118
    parameters at the defaults; `skip_connections` / `num_color_layers` are stored and ignored
    exactly like the reference).  forward runs the launch plan of nfs_b200/mlp_g3.py: every dense
    layer on tcgen05, encodings / gate / activations fused into the operand builds and epilogues.
    dino_dim=0 with dino_features of shape (N,0) (or None) is the feature-less variant."""

    def __init__(self, pos_freq=10, dir_freq=4, dino_dim=64, hidden_dim=256, num_density_layers=8,
                 num_color_layers=2, skip_connections=[4]):
        super().__init__()
        self.pos_encoder = PositionalEncoding(pos_freq)
        self.dir_encoder = PositionalEncoding(dir_freq)
        self.pos_dim = self.pos_encoder.get_output_dim(3)
        self.dir_dim = self.dir_encoder.get_output_dim(3)
        self.dino_dim = dino_dim
        try:
            from lora_dino import NeRFDINOFusion      # the reference's spelling (nerf_mlp.py:110)
        except ImportError:
            from models.lora_dino import NeRFDINOFusion
        self.dino_fusion = NeRFDINOFusion(pos_dim=self.pos_dim, dino_dim=dino_dim, hidden_dim=hidden_dim)
        self.density_mlp = DensityMLP(input_dim=hidden_dim, hidden_dim=hidden_dim, num_layers=num_density_layers)
        self.color_mlp = ColorMLP(feature_dim=hidden_dim, dir_dim=self.dir_dim, hidden_dim=hidden_dim // 2)
        self.skip_connections = skip_connections
        self._plan = None

    def _get_plan(self):
        if self._plan is None:
            self._plan = _g3.G3Plan(self)
        return self._plan

    def forward(self, positions, directions, dino_features):
        """positions (N,3), directions (N,3), dino_features (N,dino_dim) -> rgb (N,3), density (N,1)."""
        return _g3.g3_forward(self._get_plan(), positions, directions, dino_features)


def _train_py_model(pos_freq=10, dir_freq=4, hidden_dim=256, num_density_layers=8, use_dino=True, dino_dim=64,
                    **unused):
    """The model train.py:82-89 asks for with NeRFWithDINO's keyword arguments under the name
    NeRFMLP (reference broken as shipped): the view-dependent topology, without
    the feature branch when use_dino is False (dino_features=None is then accepted)."""
    return NeRFWithDINO(pos_freq=pos_freq, dir_freq=dir_freq, dino_dim=dino_dim if use_dino else 0,
                        hidden_dim=hidden_dim, num_density_layers=num_density_layers)


def _g1_class():
    try:
        from .nerf_model import NeRFMLP as G1
    except ImportError:
        from nerf_model import NeRFMLP as G1
    return G1


class NeRFMLP(_g1_class()):
    """Legacy spelling `from models.nerf_mlp import NeRFMLP` of the reference's older scripts
    (src/training/train_minimal.py:28 `NeRFMLP(pos_dim=63)`, train_lora.py:57
    `NeRFMLP(pos_dim=..., dino_dim=768, hidden_dim=256, n_layers=8, lora_rank=4)`).
    The class body is absent from the reference snapshot, so only the form whose meaning the call sites pin is
    served: without image features and without LoRA it is the plain 8 x 256 MLP of nerf_model.py:5-24 and returns
    (P,4) = [rgb | sigma] through the fused tcgen05 chain.  The feature-conditioned / LoRA forms raise - there is
    nothing to check them against (and no eager fallback to hide behind)."""

    def __init__(self, pos_dim=63, dino_dim=0, hidden_dim=256, n_layers=8, lora_rank=0):
        if dino_dim or lora_rank:
            raise NotImplementedError(
                "nfs_b200: models.nerf_mlp.NeRFMLP with dino_dim / lora_rank: the reference snapshot does not contain "
                "this class (SURVEY.md section 8f rank 4); use models.nerf_mlp.NeRFWithDINO for feature-conditioned models")
        super().__init__(pos_dim=pos_dim, hidden_dim=hidden_dim, n_layers=n_layers)

    def forward(self, x, dino_feat=None):
        if dino_feat is not None and dino_feat.shape[-1] != 0:
            raise NotImplementedError("nfs_b200: this legacy NeRFMLP was built without image features (dino_dim=0)")
        return super().forward(x)


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
