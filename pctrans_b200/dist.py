"""Batch sharding for the MSDeformAttn path.

The op is independent per image, so N ranks each take a contiguous slice of the batch dimension of
every tensor and run the kernels locally: no halo, no data-path collective.  The only exchange the
reference performs for this path is DDP's parameter-gradient all-reduce
(connectomics/model/build.py:89); ``allreduce_mean_`` is that step for a list of gradient tensors
(one flat bucket, NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import torch


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous block of ``n_items`` owned by ``rank``; earlier ranks get the
    remainder (sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world_size: int, rank: int):
    """Slice dim 0 (the image batch) of every tensor for this rank (contiguous views)."""
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("all sharded tensors must share the batch dimension")
    lo, hi = shard_bounds(n, world_size, rank)
    return [t[lo:hi].contiguous() for t in tensors]


def allreduce_mean_(grads: Iterable[torch.Tensor], group=None, async_op: bool = False):
    """Average gradient tensors across ranks in place through ONE flat bucket (the op-only
    benchmark has 205 600 parameters; the 6-layer encoder 4.39 M = 16.8 MiB).  Returns the work
    handle and a ``finish()`` callable when ``async_op`` is set."""
    import torch.distributed as dist

    grads = [g for g in grads if g is not None]
    if not grads:
        return None
    world = dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if async_op:
            work.wait()
        flat.div_(world)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    if async_op:
        return finish
    finish()
    return None
