"""CPU: pin the oracle (oracle/) against golden vectors made by the reference's own function.

The fixtures come from tests/golden/make_golden.py, which runs the reference's
ms_deform_attn_core_pytorch (OPS/functions/ms_deform_attn_func.py:52-72) and its autograd
backward.  Both restatements (C closed form, torch grid_sample port) must reproduce them.
"""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import msda_oracle as oracle
from pctrans_b200 import synth


def _load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def _full_cases(golden_dir):
    return sorted(p for p in glob.glob(os.path.join(golden_dir, "*.npz"))
                  if "cfg1" not in os.path.basename(p))


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_golden_files_present(golden_dir):
    names = {os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    for d in (2, 30, 32, 64, 71):
        assert f"kat_opstest_D{d}.npz" in names
    assert {"ragged_O_f64.npz", "ragged_O_f32.npz", "cfg1_R_f32.npz"} <= names


def test_c_oracle_matches_reference_golden(golden_dir):
    for path in _full_cases(golden_dir):
        g = _load(path)
        f32 = g["value"].dtype == np.float32
        out = oracle.forward_c(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"])
        gv, gl, ga = oracle.backward_c(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"],
                                       g["grad_out"])
        # fp64 fixtures: agreement to rounding; fp32 fixtures: the reference ran in fp32
        tol = 2e-6 if f32 else 1e-12
        assert _rel(out, g["out"]) < tol, path
        assert _rel(gv, g["grad_value"]) < tol, path
        assert _rel(gl, g["grad_loc"]) < (2e-5 if f32 else 1e-11), path
        assert _rel(ga, g["grad_attn"]) < tol, path


def test_c_oracle_f32_arithmetic(golden_dir):
    g = _load(os.path.join(golden_dir, "ragged_O_f32.npz"))
    out = oracle.forward_c(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], dtype=np.float32)
    gv, gl, ga = oracle.backward_c(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"],
                                   g["grad_out"], dtype=np.float32)
    assert out.dtype == np.float32
    assert _rel(out, g["out"]) < 1e-5
    assert _rel(gv, g["grad_value"]) < 1e-5
    assert _rel(gl, g["grad_loc"]) < 1e-4
    assert _rel(ga, g["grad_attn"]) < 1e-5


def test_torch_port_matches_reference_golden(golden_dir):
    for path in _full_cases(golden_dir):
        g = _load(path)
        t = {k: torch.from_numpy(v) for k, v in g.items()}
        value = t["value"].clone().requires_grad_(True)
        loc = t["loc"].clone().requires_grad_(True)
        attn = t["attn"].clone().requires_grad_(True)
        out = oracle.core_grid_sample(value, t["shapes"], loc, attn)
        out.backward(t["grad_out"])
        f32 = g["value"].dtype == np.float32
        tol = 1e-5 if f32 else 1e-12
        assert out.shape == t["out"].shape
        assert _rel(out.detach().numpy(), g["out"]) < tol, path
        assert _rel(value.grad.numpy(), g["grad_value"]) < tol, path
        assert _rel(loc.grad.numpy(), g["grad_loc"]) < (1e-4 if f32 else 1e-11), path
        assert _rel(attn.grad.numpy(), g["grad_attn"]) < tol, path


def test_cfg1_subsampled_golden(golden_dir):
    """BASELINE config 1: regenerate the inputs, check their bits, compare sub-sampled results."""
    g = _load(os.path.join(golden_dir, "cfg1_R_f32.npz"))
    value, shapes, lsi, loc, attn = synth.make_inputs(2, synth.CVPPP_512, M=8, D=32, P=4,
                                                      regime="R", seed=0)
    grad_out = torch.randn(2, value.shape[1], 256, generator=torch.Generator().manual_seed(1))
    h = hashlib.sha256()
    for t in (value, loc, attn, grad_out):
        h.update(t.contiguous().numpy().tobytes())
    assert h.digest() == g["input_sha256"].tobytes(), (
        "torch's CPU RNG stream no longer reproduces the inputs tests/golden/cfg1_R_f32.npz was made "
        "from: regenerate the fixture with tests/golden/make_golden.py in the authoring container")
    stride = int(g["stride"])
    out = oracle.forward_c(value, shapes, lsi, loc, attn)
    gv, gl, ga = oracle.backward_c(value, shapes, lsi, loc, attn, grad_out)
    for name, mine in (("out", out), ("grad_value", gv), ("grad_loc", gl), ("grad_attn", ga)):
        ref = g[name]
        err = np.abs(mine.reshape(-1)[::stride] - ref).max() / float(g[name + "_absmax"])
        assert err < (1e-4 if name == "grad_loc" else 1e-5), (name, err)


def test_out_of_range_points_have_zero_grads():
    value, shapes, lsi, loc, attn = synth.make_inputs(1, ((4, 5), (2, 3)), M=2, D=4, P=2,
                                                      regime="U", seed=5, Lq=7)
    loc[..., 0] = -3.0      # every point outside every map (cuh:293)
    out = oracle.forward_c(value, shapes, lsi, loc, attn)
    gv, gl, ga = oracle.backward_c(value, shapes, lsi, loc, attn, np.ones_like(out))
    assert not out.any() and not gv.any() and not gl.any() and not ga.any()


def test_linearity_in_value_and_attn():
    value, shapes, lsi, loc, attn = synth.make_inputs(2, ((6, 4), (3, 2)), M=2, D=8, P=2,
                                                      regime="O", seed=9, Lq=13)
    value, attn = value.double(), attn.double()
    a = oracle.forward_c(value, shapes, lsi, loc, attn)
    b = oracle.forward_c(2.5 * value, shapes, lsi, loc, attn)
    c = oracle.forward_c(value, shapes, lsi, loc, 0.5 * attn)
    np.testing.assert_allclose(b, 2.5 * a, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c, 0.5 * a, rtol=1e-12, atol=1e-14)


def test_empty_query_set():
    value, shapes, lsi, loc, attn = synth.make_inputs(1, ((3, 3),), M=1, D=4, P=1,
                                                      regime="U", seed=1, Lq=0)
    out = oracle.forward_c(value, shapes, lsi, loc, attn)
    assert out.shape == (1, 0, 4)


def test_port_matches_the_reference_bytecode_when_built():
    """oracle/_ref/ holds the reference's own ms_deform_attn_core_pytorch (byte-compiled from
    /root/reference by oracle/build_ref.py).  When it is there, the in-repo port and the C
    restatement must agree with it; when it is not (a checkout that never saw the reference), the
    golden vectors above are what pins the oracle."""
    ref_core = oracle.reference_core()
    if ref_core is None:
        pytest.skip("oracle/_ref not built in this checkout (python oracle/build_ref.py)")
    value, shapes, lsi, loc, attn = synth.make_inputs(2, ((9, 7), (5, 4), (3, 2)), M=4, D=8, P=3,
                                                      regime="O", seed=21, Lq=33)
    grad_out = torch.randn(2, 33, 32, generator=torch.Generator().manual_seed(22))
    outs = []
    for fn in (lambda v, lo, at: ref_core(v, [tuple(r) for r in shapes.tolist()], lo, at),
               lambda v, lo, at: oracle.core_grid_sample(v, shapes, lo, at)):
        v, lo, at = (t.clone().requires_grad_(True) for t in (value, loc, attn))
        out = fn(v, lo, at)
        out.backward(grad_out)
        outs.append((out.detach(), v.grad, lo.grad, at.grad))
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    c_out = oracle.forward_c(value, shapes, lsi, loc, attn)
    assert np.abs(c_out - outs[0][0].numpy()).max() <= 1e-5 * max(np.abs(c_out).max(), 1e-30)
