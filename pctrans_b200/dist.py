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


class BucketedGradReducer:
    """The reference trains this path under DistributedDataParallel (model/build.py:89), which
    all-reduces parameter gradients in buckets WHILE the backward pass is still running.  This is
    that exchange for the encoder around the op, without wrapping the module: parameters are
    grouped into buckets in reverse registration order (the order backward produces gradients in);
    when the last gradient of a bucket has been accumulated, the bucket is flattened and
    all-reduced -- on a side CUDA stream for NCCL, as an async work item for gloo -- while backward
    continues with the earlier layers.  ``finish()`` (call it where the optimizer would read the
    gradients) waits for the buckets, averages and writes the results back into ``p.grad``.
    """

    def __init__(self, params, bucket_bytes: int = 4 << 20, group=None):
        import torch.distributed as dist

        self._dist, self._group = dist, group
        self.world = dist.get_world_size(group)
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # lists of parameters
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {p: i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._inflight = {}        # bucket index -> (flat tensor, work handle or None)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._flat = [torch.empty(sum(p.numel() for p in b), dtype=b[0].dtype, device=dev)
                      for b in self.buckets]
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _on_grad(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        flat, bucket = self._flat[i], self.buckets[i]
        if self._side is not None:
            # the side stream picks up after the kernels that produced these gradients
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                torch.cat([p.grad.reshape(-1) for p in bucket], out=flat)
                self._dist.all_reduce(flat, group=self._group)
            self._inflight[i] = (flat, None)
        else:
            torch.cat([p.grad.reshape(-1) for p in bucket], out=flat)
            self._inflight[i] = (flat, self._dist.all_reduce(flat, group=self._group, async_op=True))

    def finish(self):
        """Wait for every bucket, average, copy back into the ``.grad`` tensors, re-arm."""
        for i, n in enumerate(self._pending):       # parameters that received no gradient this step
            if n != 0 and n != len(self.buckets[i]):
                raise RuntimeError("BucketedGradReducer: a bucket was only partly filled; every "
                                   "parameter of a bucket must take part in the backward pass")
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        for i, (flat, work) in sorted(self._inflight.items()):
            if work is not None:
                work.wait()
            flat.div_(self.world)
            off = 0
            for p in self.buckets[i]:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
