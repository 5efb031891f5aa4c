"""Data-parallel plumbing for the ray-sharded render / train path.

Rays are independent units: every kernel of the path is per-ray, so N GPUs need no data-path
collective.  The only exchange of a training step is one sum-allreduce of the flat fp32
gradient buffer of the small MLP (1.9 MB for nerf_model.NeRFMLP) - NCCL over NVLink on the GPU
box, gloo in the CPU tests of this host-side logic.
"""
import os

import torch


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items, rank, world_size):
    """Contiguous [start, stop) slice of n_items for `rank`: sizes differ by at most one, every
    item is owned exactly once (full-frame render: rays split 1/G per GPU, SURVEY.md 8d cfg 5)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(perm, batch_start, batch_size, rank, world_size):
    """Rank's share of the global batch perm[batch_start : batch_start+batch_size] - every rank
    walks the same permutation (same seed) so that G GPUs see exactly the rays 1 GPU would."""
    stop = min(batch_start + batch_size, perm.numel())
    lo, hi = shard_range(stop - batch_start, rank, world_size)
    return perm[batch_start + lo: batch_start + hi]


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of the flat gradient buffer; no-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class NcclAllreduce:
    """Sum of the flat gradient over the ranks as one NCCL call (eager: GraphedStep then runs
    [graph: render + backward] -> this -> [graph: Adam]).  Fallback of PeerExchange."""
    describe = "nccl sum-allreduce of the flat fp32 gradient, eager between two CUDA graphs"
    in_graph = False

    def __call__(self, flat):
        return allreduce_sum_(flat)


class PeerExchange:
    """Gradient exchange fused into the optimizer step (csrc/dp.cu): every rank's flat fp32 gradient lives in a
    CUDA-IPC-shared buffer; nfs_dp_adam_step sums the buffers of all ranks over NVLink while it applies Adam, with
    flag-based hand-shakes inside the kernels - no NCCL call and no graph split on the step's path.

    Construct it right after the FusedAdam and before anything captures optimizer.grad (StepSession, CUDA graphs):
    it re-homes optimizer.grad into the shared buffer.  All ranks of `group` must construct it collectively."""
    in_graph = True

    def __init__(self, optimizer, group=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self.opt, self.group = optimizer, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("PeerExchange: at most 8 ranks (one NVLink / NVSwitch box)")
        dev = optimizer.flat.device
        self.n = int(optimizer.flat.numel())
        base = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.call("nfs_dp_alloc", self.n, ctypes.byref(base))
            handle = (ctypes.c_ubyte * 64)()
            _lib.call("nfs_dp_ipc_export", base, handle)
        self.base = base.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.bases = (ctypes.c_void_p * self.world)()
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases[r] = self.base
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            out = ctypes.c_void_p()
            with torch.cuda.device(dev):
                _lib.call("nfs_dp_ipc_open", buf, ctypes.byref(out))
            self.bases[r] = out.value
            self._opened.append(out.value)
        self.grad = _tensor_from_ptr(self.base, self.n, dev)
        self.grad.zero_()
        optimizer.grad = self.grad                       # weight-gradient kernels now accumulate into the shared buffer
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        self.cta_counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self.describe = ("fused into the Adam kernel: every rank reads its %d peers' flat fp32 gradients (%d floats) over "
                         "NVLink peer memory (CUDA IPC), flag hand-shake in-kernel, inside the step's CUDA graph"
                         % (self.world - 1, self.n))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                        # every rank has mapped every buffer before the first kernel

    def wait_readers(self):
        """Before this rank's gradient buffer is written again: all peers have finished reading the last exchange."""
        from . import _lib
        from .ops import _stream
        with torch.cuda.device(self.grad.device):
            _lib.call("nfs_dp_wait_readers", self.bases, self.world, self.rank, self.n, _lib.ptr(self.epoch), _stream())

    def fused_step(self, grad_scale=1.0):
        """Sum over the ranks + Adam update, one kernel (optimizer.step() of the data-parallel step)."""
        from . import _lib, mlp
        from .ops import _stream
        o = self.opt
        o.sync_lr()
        with torch.cuda.device(o.flat.device):
            _lib.call("nfs_dp_adam_step", _lib.ptr(o.flat), self.bases, self.world, self.rank, _lib.ptr(o.exp_avg),
                      _lib.ptr(o.exp_avg_sq), self.n, float(o.betas[0]), float(o.betas[1]), float(o.eps),
                      float(o.weight_decay), _lib.ptr(o._step_dev), _lib.ptr(o._state), float(grad_scale),
                      int(bool(o.decoupled)), _lib.ptr(self.epoch), _lib.ptr(self.cta_counter), _stream())
        mlp.bump_weight_epoch()

    def __call__(self, flat):
        raise RuntimeError("PeerExchange has no stand-alone all-reduce: use wait_readers() / fused_step()")


class _RawCuda:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _tensor_from_ptr(ptr, n, device):
    """fp32 tensor of n elements over raw device memory (memory owned by the caller)."""
    return torch.as_tensor(_RawCuda(ptr, n), device=device)


def make_allreduce(optimizer, group=None):
    """The gradient exchange of a data-parallel training step for `optimizer` (FusedAdam): the peer-memory exchange fused
    into the Adam kernel when every rank can map every other rank's buffer (one box, CUDA IPC), else one NCCL
    all-reduce.  NFS_DP_EXCHANGE=nccl forces the fallback.  Collective over the ranks of `group`."""
    import os
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    ok, ex, err = 1, None, ""
    if os.environ.get("NFS_DP_EXCHANGE", "p2p") == "nccl" or not hasattr(optimizer, "flat") or dist.get_world_size(group) > 8:
        ok = 0
    else:
        try:
            ex = PeerExchange(optimizer, group)
        except Exception as e:            # e.g. IPC not permitted in this container: every rank must agree on the fallback
            ok, err = 0, repr(e)[:200]
    flag = torch.tensor([ok], device=optimizer.flat.device if hasattr(optimizer, "flat") else "cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return ex
    nccl = NcclAllreduce()
    if err:
        nccl.describe += " (peer exchange unavailable: %s)" % err
    return nccl


def loss_scale(local_count, global_count):
    """A mean over the global batch is the sum over ranks of local_sum / global_count: scale the
    local mean loss by local_count / global_count before backward, then sum-allreduce."""
    return float(local_count) / float(global_count)


def gather_rows(local, counts, group=None):
    """All-gather row blocks of unequal height (rendered rgb tiles) in rank order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)
