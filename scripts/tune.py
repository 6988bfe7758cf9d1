#!/usr/bin/env python
"""Tuning / profiling harness (GPU box): times the raw C-ABI forward and backward calls with
CUDA events for a list of launch-shape flags, dtypes and location regimes.

    python scripts/tune.py [--batch 16] [--pyramid cvppp512] [--iters 30] [--flags 0,0x10800,...]
                           [--dtypes f32,bf16] [--regimes R,U] [--d-model 256] [--only fwd|bwd]

flags: bits 8..15 = threads/32 of the persistent CTA, bits 16..19 = CTAs per SM (0 = default).
Under ncu use --iters 1 --warm 1.
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pctrans_b200 import _cabi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--pyramid", default="CVPPP_512")
    ap.add_argument("--d-model", type=int, default=256)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--warm", type=int, default=5)
    ap.add_argument("--flags", default="0")
    ap.add_argument("--dtypes", default="f32")
    ap.add_argument("--regimes", default="R")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pyramid = getattr(synth, args.pyramid)
    D = args.d_model // 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for regime in args.regimes.split(","):
        value, shapes, lsi, loc, attn = synth.make_inputs(args.batch, pyramid, M=8, D=D, P=4,
                                                          regime=regime, seed=0)
        S = value.shape[1]
        nq = args.batch * S
        gout = torch.randn(args.batch, S, 8 * D, generator=torch.Generator().manual_seed(1))
        sh, ls, lo, at = shapes.to(dev), lsi.to(dev), loc.to(dev), attn.to(dev)
        for dn in args.dtypes.split(","):
            dt = torch.float32 if dn == "f32" else torch.bfloat16
            v, g = value.to(dev, dt), gout.to(dev, dt)
            vb = 4 if dn == "f32" else 2
            bf = 2 * 8 * D * vb + 8 * 12 * 12
            bb = 3 * 8 * D * vb + 2 * 8 * 12 * 12
            out = torch.empty(args.batch, S, 8 * D, dtype=dt, device=dev)
            grads = (torch.empty_like(v), torch.empty_like(lo), torch.empty_like(at))
            ws = torch.empty(v.numel() * 4 if dn == "bf16" else 0, dtype=torch.uint8, device=dev)
            for fs in args.flags.split(","):
                flags = int(fs, 0)
                res = {}
                for name, fn in (("fwd", lambda: _cabi.forward(v, sh, ls, lo, at, flags=flags, out=out)),
                                 ("bwd", lambda: _cabi.backward(v, sh, ls, lo, at, g, flags=flags,
                                                                out=grads, workspace=ws))):
                    if args.only and args.only != name:
                        continue
                    for _ in range(args.warm):
                        fn()
                    ts = []
                    for _ in range(args.iters):
                        flush.zero_()          # evict L2 between timed launches
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        fn()
                        e1.record()
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    res[name] = statistics.median(ts)
                msg = f"{regime} {dn} D={D} N={args.batch} flags={flags:#x}:"
                if "fwd" in res:
                    msg += f" fwd {res['fwd'] * 1e3:7.1f} us {nq / res['fwd'] / 1e3:7.1f} Mq/s {nq * bf / res['fwd'] / 1e6:6.0f} GB/s |"
                if "bwd" in res:
                    msg += f" bwd {res['bwd'] * 1e3:7.1f} us {nq / res['bwd'] / 1e3:7.1f} Mq/s {nq * bb / res['bwd'] / 1e6:6.0f} GB/s"
                print(msg, flush=True)


if __name__ == "__main__":
    main()
