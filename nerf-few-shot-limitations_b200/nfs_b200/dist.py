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
    [graph: render + backward] -> this -> [graph: Adam])."""
    describe = "nccl sum-allreduce of the flat fp32 gradient, eager between two CUDA graphs"
    in_graph = False

    def __call__(self, flat):
        return allreduce_sum_(flat)


def make_allreduce(flat):
    """The gradient exchange of a data-parallel training step for the flat fp32 gradient buffer `flat`."""
    return NcclAllreduce()


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
