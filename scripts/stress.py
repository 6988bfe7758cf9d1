#!/usr/bin/env python
"""Randomised stress run (GPU box): many random shapes / dtypes / regimes / entry points against
the oracle.  `python scripts/stress.py [n_cases] [seed]`; prints the worst errors and exits 1 on
any violation.  Benchmark-side tool: the oracle is test infrastructure."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import msda_oracle as oracle
from pctrans_b200 import _cabi, synth
from pctrans_b200.functions import MSDeformAttnFusedFunction

dev = "cuda:0"
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dims = [1, 2, 3, 4, 8, 12, 16, 20, 32, 48, 64, 128, 256]


def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = np.asarray(b, np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-30))


worst = {}
fails = 0
t0 = time.time()
for case in range(n_cases):
    L = rng.randint(1, 5)
    shapes = tuple((rng.randint(1, 20), rng.randint(1, 20)) for _ in range(L))
    M, D, P = rng.choice([1, 2, 4, 8, 16]), rng.choice(dims), rng.randint(1, 8)
    N, Lq = rng.randint(1, 4), rng.randint(1, 300)
    dtype = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float64])
    regime = rng.choice(["U", "O", "O"])
    mode = rng.choice(["plain", "plain", "det", "fused2", "fused4", "binned", "binned"])
    value, sh, lsi, loc, attn = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime=regime, seed=case, Lq=Lq)
    loc = synth.nudge_off_grid(loc, shapes, eps=2e-3)
    g = torch.randn(N, Lq, M * D, generator=torch.Generator().manual_seed(case + 7)) * rng.choice([1e-3, 1.0, 1e3])
    if dtype == torch.bfloat16:
        value, g = value.bfloat16().float(), g.bfloat16().float()
    ldt = torch.float64 if dtype == torch.float64 else torch.float32
    d = lambda t, dt=None: t.to(dev, dt) if dt else t.to(dev)
    v, go = d(value, dtype), d(g, dtype)
    ftol, btol = {torch.float32: (1e-5, 1e-4), torch.bfloat16: (2e-2, 2e-2), torch.float64: (1e-11, 1e-10)}[dtype]
    vec = dtype != torch.float64 and _cabi.kernel_family(D, L, P, dtype) == 1 and _cabi.kernel_family(D, L, P, dtype, True) == 1
    try:
        if mode.startswith("fused") and vec:
            rd = int(mode[-1])
            gen = torch.Generator().manual_seed(case + 99)
            norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
            if rd == 2:
                ref = torch.rand(N, Lq, L, 2, generator=gen)
                off = (loc - ref[:, :, None, :, None, :]) * norm[None, None, None, :, None, :]
                dloc = (1.0 / norm.double())[None, None, None, :, None, :].expand_as(off)
                loc64 = ref.double()[:, :, None, :, None, :] + off.double() * dloc
            else:
                ref = torch.cat([torch.rand(N, Lq, L, 2, generator=gen), torch.rand(N, Lq, L, 2, generator=gen) * 0.5 + 0.1], -1)
                off = (loc - ref[:, :, None, :, None, :2]) * P / (ref[:, :, None, :, None, 2:] * 0.5)
                dloc = (ref.double()[:, :, None, :, None, 2:] * 0.5 / P).expand_as(off)
                loc64 = ref.double()[:, :, None, :, None, :2] + off.double() * dloc
            logits = torch.randn(N, Lq, M, L * P, generator=gen) * 2
            a64 = torch.softmax(logits.double(), -1).view(N, Lq, M, L, P)
            out = _cabi.fused_forward(v, d(sh), d(lsi), d(ref.contiguous()), d(off.contiguous()), d(logits))
            gv, goff, glog = _cabi.fused_backward(v, d(sh), d(lsi), d(ref.contiguous()), d(off.contiguous()), d(logits), go)
            r_out = oracle.forward_c(value, sh, lsi, loc64, a64)
            r_gv, r_gl, r_ga = oracle.backward_c(value, sh, lsi, loc64, a64, g)
            r_goff = torch.from_numpy(r_gl) * dloc
            ga_t = torch.from_numpy(r_ga).view(N, Lq, M, L * P); a = a64.view(N, Lq, M, L * P)
            r_glog = a * (ga_t - (a * ga_t).sum(-1, keepdim=True))
            errs = dict(out=rel(out, r_out) / ftol, gv=rel(gv, r_gv) / btol, gl=rel(goff, r_goff.numpy()) / btol, ga=rel(glog, r_glog.numpy()) / btol)
        else:
            det_ok = vec and D <= 128   # deterministic kernels exist for 4-channel lanes (<= 32 lanes) only
            flags = _cabi.FLAG_DETERMINISTIC if (mode == "det" and det_ok) else 0
            if mode == "binned":   # the cell-binned backward wherever the shape fits it
                flags = _cabi.FLAG_BWD_BINNED
            lo, at = d(loc, ldt), d(attn, ldt)
            out = _cabi.forward(v, d(sh), d(lsi), lo, at)
            gv, gl, ga = _cabi.backward(v, d(sh), d(lsi), lo, at, go, flags=flags)
            r_out = oracle.forward_c(value, sh, lsi, loc, attn)
            r_gv, r_gl, r_ga = oracle.backward_c(value, sh, lsi, loc, attn, g)
            errs = dict(out=rel(out, r_out) / ftol, gv=rel(gv, r_gv) / btol, gl=rel(gl, r_gl) / btol, ga=rel(ga, r_ga) / btol)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"case {case} EXCEPTION {type(e).__name__}: {e} :: shapes={shapes} M={M} D={D} P={P} N={N} Lq={Lq} {dtype} {regime} {mode}")
        fails += 1
        continue
    bad = {k: v for k, v in errs.items() if not (v <= 1.0)}
    for k, v in errs.items():
        key = (str(dtype).split(".")[-1], mode if vec or mode == "plain" else "plain", k)
        worst[key] = max(worst.get(key, 0.0), v)
    if bad:
        fails += 1
        print(f"case {case} FAIL {bad} :: shapes={shapes} M={M} D={D} P={P} N={N} Lq={Lq} {dtype} {regime} {mode} vec={vec}")
print(f"{n_cases} cases in {time.time() - t0:.1f}s, failures: {fails}")
for k in sorted(worst):
    print("  worst err / tolerance", k, f"{worst[k]:.3f}")
sys.exit(1 if fails else 0)
