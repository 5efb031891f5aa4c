"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/nfs_b200.h declares, the ctypes prototypes cover exactly those symbols, and the
product path refuses CPU tensors (no fallback).  No kernel is launched here."""
import ctypes
import os
import subprocess

import pytest
import torch


def test_library_is_built_and_exports_header_symbols():
    from nfs_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py build` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 8
    for name in declared:
        assert hasattr(lib, name), "libnfs_b200.so lacks %s declared in include/nfs_b200.h" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T nfs_" in l)
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_abi_version_and_error_string():
    from nfs_b200 import _lib
    lib = _lib.load()
    assert lib.nfs_abi_version() == 7
    assert isinstance(_lib.last_error(), str)
    assert _lib.launch_count() >= 0


def test_argument_errors_do_not_need_a_gpu():
    """Negative sizes / null pointers are rejected before any CUDA call."""
    from nfs_b200 import _lib
    lib = _lib.load()
    rc = lib.nfs_composite_fwd(None, None, None, None, None, 0.0, 10, 64, 0, 0, None, None, None, None)
    assert rc == -1 and "nfs_composite_fwd" in _lib.last_error()
    rc = lib.nfs_composite_fwd(None, None, None, None, None, 0.0, -1, 64, 0, 0, None, None, None, None)
    assert rc == -1
    assert lib.nfs_composite_fwd(None, None, None, None, None, 0.0, 0, 64, 0, 0, None, None, None, None) == 0
    rc = lib.nfs_sample_hierarchical(None, None, None, None, None, 0, None, 4, 5000, 10, None, None, None, None, None, None)
    assert rc < 0
    with pytest.raises(RuntimeError, match="nfs_posenc_fwd"):
        _lib.call("nfs_posenc_fwd", None, None, 5, 3, 10, 1, None, None)


def test_product_path_has_no_cpu_fallback():
    from models.nerf_mlp import PositionalEncoding, VolumeRenderer
    from models.volume_renderer import volume_render_radiance
    from utils.ray_utils import hierarchical_sampling, sample_points_along_rays
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VolumeRenderer()(torch.rand(2, 4, 3), torch.rand(2, 4, 1), torch.rand(2, 4), torch.rand(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        volume_render_radiance(torch.rand(2, 4, 4), torch.rand(2, 4), torch.rand(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PositionalEncoding(4)(torch.rand(5, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sample_points_along_rays(torch.rand(5, 3), torch.rand(5, 3), 2.0, 6.0, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hierarchical_sampling(torch.rand(5, 3), torch.rand(5, 3), torch.rand(5, 8), torch.rand(5, 7), 4)
    # the 8f additions: ray generation, loss epilogue, feature gather, MLPs, fused optimizer
    from models.nerf_mlp import NeRFWithDINO
    from models.nerf_model import NeRFMLP
    from models.ray_sampler import get_rays
    from nfs_b200 import ops
    from nfs_b200.optim import FusedAdam
    from utils.ray_utils import get_rays as get_rays_u, project_points_to_image
    for fn in (get_rays, get_rays_u):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(4, 5, 3.0, torch.eye(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.composite_loss(torch.rand(2, 4, 3), torch.rand(2, 4), torch.rand(2, 4), torch.rand(2, 3), torch.rand(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        project_points_to_image(torch.rand(6, 3), torch.eye(4), 10.0, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NeRFMLP()(torch.rand(3, 63))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NeRFWithDINO()(torch.rand(3, 3), torch.rand(3, 3), torch.rand(3, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdam(NeRFMLP().parameters())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.wgrad_multi([dict(u=torch.zeros(64, 128, dtype=torch.bfloat16), v=torch.zeros(64, 64, dtype=torch.bfloat16),
                              dw=torch.zeros(64, 128), ld_m=1, ld_n=128)])


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerf-few-shot-limitations_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "nerf_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_both_import_spellings():
    """`from models.nerf_mlp import X` (train.py:19-25) and bare `from nerf_mlp import X`
    (train_multiscale.py:15-17) both resolve to the drop-in."""
    import sys
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerf-few-shot-limitations_b200")
    code = ("import sys; sys.path[:0]=[%r, %r]; import nerf_mlp, ray_utils, volume_renderer, positional_encoding, "
            "ray_sampler; print(nerf_mlp.VolumeRenderer.__name__, ray_utils.hierarchical_sampling.__name__)"
            % (os.path.join(root, "models"), os.path.join(root, "utils")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "VolumeRenderer hierarchical_sampling" in out.stdout
