"""Makes `import nfs_b200` work when a models/ or utils/ module is imported by its bare
name (the reference's train_multiscale.py:15-17 / evaluate.py:9-10 put src/models and
src/utils themselves on sys.path and do `from nerf_mlp import ...`)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
