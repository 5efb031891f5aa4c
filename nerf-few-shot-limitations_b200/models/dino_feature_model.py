"""Drop-in for the hot-path part of /root/reference/src/models/dino_feature_model.py:
NeRFDINOFusion (:150-197).  The rest of that file (LoRALinear, SpatialDINOFeatures: a HuggingFace
Dinov2 backbone that runs once per view under no_grad and needs downloaded weights) is outside
the render hot path - SURVEY.md 2.1 rows 8-11."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch
import torch.nn as nn

from nfs_b200 import mlp_g3 as _g3
from nfs_b200 import ops as _ops


def sample_features_at_points(features, points_2d):
    """The body of SpatialDINOFeatures.sample_features_at_points (dino_feature_model.py:114-148): bilinear
    F.grid_sample (zeros padding, align_corners=False) of a (1, Hp, Wp, C) feature map at normalised points
    (N,2) -> (N,C), as one kernel.  The class itself (a HuggingFace Dinov2 backbone, out of scope) can
    forward its method here."""
    return _ops.sample_features(features, points_2d)


class NeRFDINOFusion(nn.Module):
    """Two-layer fusion MLP applied twice around a 2-way softmax gate, then a linear projection
    (dino_feature_model.py:150-197; identical copy in lora_dino.py:146-193).  Same sub-module names:
    `fusion.{0,2}`, `attention.{0,2}`, `output_proj`.

    Inside NeRFWithDINO the fused launch plan of nfs_b200/mlp_g3.py runs instead of this forward.
    Called on its own (already-encoded fp32 inputs), every Linear runs on the tcgen05 kernels and
    the concatenation / gate product between them is torch glue on the same CUDA tensors."""

    def __init__(self, pos_dim, dino_dim, hidden_dim=256):
        super().__init__()
        self.pos_dim = pos_dim
        self.dino_dim = dino_dim
        self.fusion = nn.Sequential(nn.Linear(pos_dim + dino_dim, hidden_dim), nn.ReLU(inplace=True),
                                    nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True))
        self.attention = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 4), nn.ReLU(inplace=True),
                                       nn.Linear(hidden_dim // 4, 2), nn.Softmax(dim=-1))
        self.output_proj = nn.Linear(hidden_dim, hidden_dim)

    def _fuse(self, c):
        return _g3.dense(self, self.fusion[2], _g3.dense(self, self.fusion[0], c, "relu"), "relu")

    def forward(self, pos_encoding, dino_features):
        fused = self._fuse(torch.cat([pos_encoding, dino_features], dim=-1))
        logits = _g3.dense(self, self.attention[2], _g3.dense(self, self.attention[0], fused, "relu"))
        weights = torch.softmax(logits, dim=-1)
        final = self._fuse(torch.cat([pos_encoding * weights[:, 0:1], dino_features * weights[:, 1:2]], dim=-1))
        return _g3.dense(self, self.output_proj, final)
