#!/usr/bin/env python
"""Kernel-vs-kernel baseline (GPU box): the reference's own Deformable-DETR CUDA kernels, built
for sm_100a by baseline/build_legacy.py into the git-ignored baseline/_ref/, timed on the same
inputs as the new kernels (SURVEY.md 8d baseline iii).  Also times the reference's PyTorch
grid_sample formulation on the GPU (baseline ii).  Prints one JSON line.
"""
import argparse
import importlib.machinery
import importlib.util
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _gridsample import grid_sample_core  # noqa: E402
from pctrans_b200 import _cabi, synth  # noqa: E402


def load_legacy():
    path = os.path.join(ROOT, "baseline", "_ref", "MultiScaleDeformableAttention.so")
    if not os.path.exists(path):
        return None
    loader = importlib.machinery.ExtensionFileLoader("MultiScaleDeformableAttention", path)
    spec = importlib.util.spec_from_loader("MultiScaleDeformableAttention", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--pyramid", default="CVPPP_512")
    ap.add_argument("--d-model", type=int, default=256)
    ap.add_argument("--regime", default="R")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    D = args.d_model // 8
    value, shapes, lsi, loc, attn = synth.make_inputs(args.batch, getattr(synth, args.pyramid), M=8,
                                                      D=D, P=4, regime=args.regime, seed=0)
    S = value.shape[1]
    gout = torch.randn(args.batch, S, 8 * D, generator=torch.Generator().manual_seed(1))
    v, sh, ls, lo, at, g = (t.to(dev) for t in (value, shapes, lsi, loc, attn, gout))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nq = args.batch * S
    res = {"bench": "legacy_vs_new", "batch": args.batch, "pyramid": args.pyramid, "head_dim": D,
           "regime": args.regime, "queries": nq}

    new_f = timeit(lambda: _cabi.forward(v, sh, ls, lo, at), args.iters, flush)
    new_b = timeit(lambda: _cabi.backward(v, sh, ls, lo, at, g), args.iters, flush)
    res["new"] = {"fwd_ms": new_f, "bwd_ms": new_b, "fwd_bwd_qps": nq / ((new_f + new_b) * 1e-3)}

    legacy = load_legacy()
    if legacy is None:
        res["legacy"] = "baseline/_ref/MultiScaleDeformableAttention.so not built"
    else:
        leg_f = timeit(lambda: legacy.ms_deform_attn_forward(v, sh, ls, lo, at, 128), args.iters, flush)
        leg_b = timeit(lambda: legacy.ms_deform_attn_backward(v, sh, ls, lo, at, g, 128), args.iters, flush)
        out_l = legacy.ms_deform_attn_forward(v, sh, ls, lo, at, 128)
        gv_l, gl_l, ga_l = legacy.ms_deform_attn_backward(v, sh, ls, lo, at, g, 128)
        out_n = _cabi.forward(v, sh, ls, lo, at)
        gv_n, gl_n, ga_n = _cabi.backward(v, sh, ls, lo, at, g)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        res["legacy"] = {"fwd_ms": leg_f, "bwd_ms": leg_b, "fwd_bwd_qps": nq / ((leg_f + leg_b) * 1e-3),
                         "max_rel_diff_new_vs_legacy": {"out": rel(out_n, out_l), "grad_value": rel(gv_n, gv_l),
                                                        "grad_loc": rel(gl_n, gl_l), "grad_attn": rel(ga_n, ga_l)}}
        res["speedup_vs_legacy"] = {"fwd": leg_f / new_f, "bwd": leg_b / new_b,
                                    "fwd_bwd": (leg_f + leg_b) / (new_f + new_b)}

    def gs_step():
        vv, ll, aa = v.clone().requires_grad_(True), lo.clone().requires_grad_(True), at.clone().requires_grad_(True)
        out = grid_sample_core(vv, shapes, ll, aa)
        out.backward(g)
    gs = timeit(gs_step, max(3, args.iters // 4), flush)
    res["gridsample_gpu"] = {"fwd_bwd_ms": gs, "fwd_bwd_qps": nq / (gs * 1e-3)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
