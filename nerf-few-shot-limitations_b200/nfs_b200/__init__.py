"""nfs_b200 - B200 (sm_100a) kernels behind the NeRF render hot path of
ANKITSANJYAL/nerf-few-shot-limitations.

Layout of the drop-in (put this directory's PARENT, i.e. nerf-few-shot-limitations_b200/,
on sys.path where the reference's callers put <reference>/src):

    models/   utils/     same module / symbol names as the reference's src/models, src/utils
    nfs_b200/            ctypes binding of include/nfs_b200.h, autograd glue, sharding helpers
    csrc/                the CUDA kernels and the C ABI

No CPU path, no eager-PyTorch path: operators raise RuntimeError when the library is not
built or the tensors are not on a CUDA device.
"""
from . import _lib  # noqa: F401
from ._lib import LIB_PATH, launch_count, load  # noqa: F401

__all__ = ["LIB_PATH", "launch_count", "load"]
