"""CPU, world_size 2, gloo: the batch-sharding logic of the multi-GPU path.

Kernels cannot run here, so the per-rank op is the oracle (test infrastructure); what is under
test is pctrans_b200.dist: shard bounds, that per-rank results concatenate to the single-process
result (the op has no cross-image term), and the gradient all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_exactly():
    from pctrans_b200.dist import shard_bounds
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import msda_oracle as oracle
    from pctrans_b200 import synth
    from pctrans_b200.dist import allreduce_mean_, shard_batch

    value, sh, lsi, loc, attn = synth.make_inputs(5, ((6, 4), (3, 2)), M=2, D=8, P=2, regime="O",
                                                  seed=0, Lq=9)
    gout = torch.randn(5, 9, 16, generator=torch.Generator().manual_seed(1))
    v, lo, at, g = shard_batch([value, loc, attn, gout], world, rank)
    out = oracle.forward_c(v, sh, lsi, lo, at)
    gv, gl, ga = oracle.backward_c(v, sh, lsi, lo, at, g)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), out=out, gv=gv, gl=gl, ga=ga)
    # DDP-style exchange: mean of per-rank "parameter gradients"
    grads = [torch.full((3, 2), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
    allreduce_mean_(grads)
    assert torch.allclose(grads[0], torch.full((3, 2), 1.5))
    assert torch.allclose(grads[1], torch.arange(4, dtype=torch.float32) * 1.5)
    fin = allreduce_mean_([torch.ones(5) * rank], async_op=True)
    fin()
    # bucketed, overlapped exchange (DDP-style hooks): equals the plain mean of the per-rank grads
    from pctrans_b200.dist import BucketedGradReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4),
                              torch.nn.ReLU(), torch.nn.Linear(4, 3))
    red = BucketedGradReducer(net.parameters(), bucket_bytes=64)
    assert len(red.buckets) >= 3
    for step in range(2):            # re-armed after finish()
        net.zero_grad(set_to_none=True)
        x = torch.randn(7, 6, generator=torch.Generator().manual_seed(10 * step + rank))
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        red.finish()
        for p, g in zip(net.parameters(), local):
            both = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(both, g)
            assert torch.allclose(p.grad, sum(both) / world, atol=1e-6)
    red.remove()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    from oracle import msda_oracle as oracle
    from pctrans_b200 import synth
    oracle.build()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    value, sh, lsi, loc, attn = synth.make_inputs(5, ((6, 4), (3, 2)), M=2, D=8, P=2, regime="O",
                                                  seed=0, Lq=9)
    gout = torch.randn(5, 9, 16, generator=torch.Generator().manual_seed(1))
    ref_out = oracle.forward_c(value, sh, lsi, loc, attn)
    ref = oracle.backward_c(value, sh, lsi, loc, attn, gout)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert np.array_equal(np.concatenate([p["out"] for p in parts]), ref_out)
    for key, full in zip(("gv", "gl", "ga"), ref):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), full)
