"""Drop-in for the hot-path part of /root/reference/src/models/lora_dino.py: nerf_mlp.py:110 imports
NeRFDINOFusion from this module (lora_dino.py:146-193, a copy of dino_feature_model.py:150-197)."""
try:
    from .dino_feature_model import NeRFDINOFusion  # noqa: F401
except ImportError:
    from dino_feature_model import NeRFDINOFusion  # noqa: F401
