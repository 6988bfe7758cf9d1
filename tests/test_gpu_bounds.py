"""GPU: memory-safety checks that stand in for compute-sanitizer (closed on the B200 pool).

1. The vector kernels are rebuilt with -DPCT_MSDA_BOUNDS_CHECK (every tap address is checked
   in-kernel and the kernel traps on a violation) and run, in a subprocess, over border-heavy
   inputs: out-of-range locations, 1-pixel maps, ragged tails, every lane layout.
2. Outputs are written into the middle of canary-filled buffers; the canaries must survive.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
from pctrans_b200 import _cabi, synth
assert _cabi.LIB_PATH.endswith("_dbg.so"), _cabi.LIB_PATH
dev = "cuda:0"
cases = [(((1, 1), (1, 3), (2, 1)), 3), (((7, 9), (4, 5), (2, 3)), 4), (((17, 22), (33, 44)), 2)]
n = nb = 0
for shapes, P in cases:
    for D, dt in ((4, torch.float32), (16, torch.float32), (32, torch.float32), (128, torch.float32),
                  (256, torch.float32), (4, torch.bfloat16), (32, torch.bfloat16), (256, torch.bfloat16)):
        for regime, lq in (("O", 37), ("U", 5), ("O", 1)):
            value, sh, lsi, loc, attn = synth.make_inputs(3, shapes, M=4, D=D, P=P, regime=regime,
                                                          seed=D + lq, Lq=lq)
            loc = (loc - 0.5) * 3.0 + 0.5          # push many points far outside the maps
            loc[0, 0, 0, 0, 0] = torch.tensor([float("nan"), 0.5])
            loc[0, 0, 1, 0, 0] = torch.tensor([1e30, -1e30])
            v = value.to(dev, dt); lo = loc.to(dev); at = attn.to(dev)
            g = torch.randn(3, lq, 4 * D, device=dev).to(dt)
            out = _cabi.forward(v, sh.to(dev), lsi.to(dev), lo, at)
            gv, gl, ga = _cabi.backward(v, sh.to(dev), lsi.to(dev), lo, at, g)
            torch.cuda.synchronize()
            assert torch.isfinite(out.float()).all() and torch.isfinite(gv.float()).all()
            assert torch.isfinite(gl).all() and torch.isfinite(ga).all()
            n += 1
            # the cell-binned backward (its index arithmetic carries device-side asserts in this build)
            S, L = v.shape[1], len(shapes)
            if _cabi.backward_plan(3, S, 4, D, L, lq, P, dt, _cabi.FLAG_BWD_BINNED) == 2:
                bv, bl, ba = _cabi.backward(v, sh.to(dev), lsi.to(dev), lo, at, g, flags=_cabi.FLAG_BWD_BINNED)
                torch.cuda.synchronize()
                assert torch.isfinite(bv.float()).all() and torch.isfinite(bl).all() and torch.isfinite(ba).all()
                assert torch.allclose(bl, gl, rtol=1e-3, atol=1e-3 * float(gl.abs().max()) + 1e-30)
                nb += 1
print("bounds-checked cases:", n)
print("binned cases:", nb)
"""


def test_vector_kernels_under_bounds_checking_build():
    from pctrans_b200 import build
    lib = build.build_library(debug=True)
    env = dict(os.environ, PCTRANS_MSDA_LIB=lib)
    proc = subprocess.run([sys.executable, "-c", _CHILD % ROOT], env=env, capture_output=True,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "bounds-checked cases: 72" in proc.stdout, proc.stdout
    assert "binned cases: 36" in proc.stdout, proc.stdout     # D in {16, 32, 128} fp32, 32 bf16


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_outputs_stay_inside_their_buffers(dtype):
    from pctrans_b200 import _cabi, synth
    dev = "cuda:0"
    value, sh, lsi, loc, attn = synth.make_inputs(3, ((9, 7), (5, 4), (2, 3)), M=8, D=32, P=4,
                                                  regime="O", seed=5, Lq=61)
    v, lo, at = value.to(dev, dtype), loc.to(dev), attn.to(dev)
    g = torch.randn(3, 61, 256, device=dev).to(dtype)
    pad = 4096

    def guarded(like):
        buf = torch.full((like.numel() + 2 * pad,), 7.0, dtype=like.dtype, device=dev)
        return buf, buf[pad:pad + like.numel()].view(like.shape)

    out_buf, out = guarded(torch.empty(3, 61, 256, dtype=dtype))
    gv_buf, gv = guarded(v)
    gl_buf, gl = guarded(lo)
    ga_buf, ga = guarded(at)
    _cabi.forward(v, sh.to(dev), lsi.to(dev), lo, at, out=out)
    _cabi.backward(v, sh.to(dev), lsi.to(dev), lo, at, g, out=(gv, gl, ga))
    torch.cuda.synchronize()
    for buf in (out_buf, gv_buf, gl_buf, ga_buf):
        assert (buf[:pad] == 7.0).all() and (buf[-pad:] == 7.0).all()
    ref = _cabi.forward(v, sh.to(dev), lsi.to(dev), lo, at)
    assert torch.equal(ref, out)
