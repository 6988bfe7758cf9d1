"""CPU: the C-ABI library loads and exports what include/pct_msda.h declares; host-side logic of
the Function / module mirrors the reference interface.  No kernels are launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    syms = set()
    for name in ("pct_msda.h", "pct_linear.h"):
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(pct_(?:msda|linear)_\w+)\s*\(", text))
    return sorted(syms)


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for s in ("pct_msda_forward", "pct_msda_backward", "pct_msda_error_string",
              "pct_msda_abi_version", "pct_msda_backward_workspace_bytes"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from pctrans_b200 import _cabi, build
    build.build_library()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported"
    assert set(_cabi.EXPORTS) | set(_cabi.LINEAR_EXPORTS) <= set(_header_symbols())


def test_abi_scalar_entry_points_without_gpu():
    from pctrans_b200 import _cabi
    lib = _cabi.load()
    assert lib.pct_msda_abi_version() == 4
    assert lib.pct_msda_error_string(0) == b"success"
    assert b"null pointer" in lib.pct_msda_error_string(1001)
    assert b"unknown" in lib.pct_msda_error_string(-5)
    assert lib.pct_msda_backward_workspace_bytes(16, 5376, 8, 32, 0) == 0
    assert lib.pct_msda_backward_workspace_bytes(16, 5376, 8, 32, 1) == 16 * 5376 * 8 * 32 * 4
    assert _cabi.kernel_family(32, 3, 4, torch.float32) == 1
    assert _cabi.kernel_family(16, 3, 4, torch.float32, backward=True) == 1
    assert _cabi.kernel_family(71, 3, 4, torch.float32) == 0
    # include/pct_linear.h: packed weight = 256-column tiles x 16-float K steps x (hi + lo) 16 KB
    assert lib.pct_linear_packed_bytes(256, 256) == 16 * 32768
    assert lib.pct_linear_packed_bytes(256, 288) == 2 * 16 * 32768
    assert lib.pct_linear_packed_bytes(250, 256) == 0 and lib.pct_linear_packed_bytes(256, 8) == 0
    # bf16 operands: 32-element K steps x one 16 KB piece
    assert lib.pct_linear_packed_bytes_bf16(256, 256) == 8 * 16384
    assert lib.pct_linear_packed_bytes_bf16(1024, 288) == 2 * 32 * 16384
    assert lib.pct_linear_packed_bytes_bf16(48, 256) == 0      # in_features % 32 != 0
    assert _cabi.kernel_family(32, 33, 4, torch.float32) == 0   # L*P > 128 sample table
    # argument validation happens before any CUDA call
    assert lib.pct_msda_forward(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 7, None) == 1003
    assert lib.pct_msda_forward(None, None, None, None, None, None, 1, 1, 0, 1, 1, 1, 1, 0, None) == 1002
    assert lib.pct_msda_forward(None, None, None, None, None, None, 0, 1, 1, 1, 1, 1, 1, 0, None) == 0


def test_function_rejects_cpu_tensors_no_fallback():
    from pctrans_b200 import synth
    from pctrans_b200.functions import MSDeformAttnFunction
    value, sh, lsi, loc, attn = synth.make_inputs(1, ((4, 4),), M=2, D=8, P=2, regime="U", seed=1, Lq=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        MSDeformAttnFunction.apply(value, sh, lsi, loc, attn, 128)


def test_module_interface_matches_reference():
    """Constructor, parameter names/shapes and initialisation of OPS/modules/ms_deform_attn.py."""
    import math
    from pctrans_b200.modules import MSDeformAttn
    with pytest.raises(ValueError):
        MSDeformAttn(d_model=100, n_heads=8)
    with pytest.warns(UserWarning):
        MSDeformAttn(d_model=24, n_levels=1, n_heads=8, n_points=1)
    m = MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4)
    names = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert names == {
        "sampling_offsets.weight": (8 * 3 * 4 * 2, 256), "sampling_offsets.bias": (8 * 3 * 4 * 2,),
        "attention_weights.weight": (8 * 3 * 4, 256), "attention_weights.bias": (8 * 3 * 4,),
        "value_proj.weight": (256, 256), "value_proj.bias": (256,),
        "output_proj.weight": (256, 256), "output_proj.bias": (256,),
    }
    assert m.im2col_step == 128
    assert not m.sampling_offsets.weight.any() and not m.attention_weights.weight.any()
    b = m.sampling_offsets.bias.detach().view(8, 3, 4, 2)
    # head 0 points along +x, point k is k+1 pixels out (mod.py:66-74)
    assert torch.allclose(b[0, :, :, 0], torch.tensor([1.0, 2.0, 3.0, 4.0]).expand(3, 4))
    assert torch.allclose(b[0, :, :, 1], torch.zeros(3, 4), atol=1e-6)
    th = 2 * math.pi * 3 / 8
    d = torch.tensor([math.cos(th), math.sin(th)])
    d = d / d.abs().max()
    assert torch.allclose(b[3, 1, 2], 3 * d, atol=1e-6)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 256), torch.zeros(1, 4, 3, 3), torch.zeros(1, 4, 256),
          torch.tensor([[2, 2], [0, 0], [0, 0]]), torch.tensor([0, 4, 4]))


def test_debug_helper_matches_oracle_on_cpu():
    import numpy as np
    from oracle import msda_oracle as oracle
    from pctrans_b200 import synth
    from pctrans_b200.functions import ms_deform_attn_core_pytorch
    value, sh, lsi, loc, attn = synth.make_inputs(2, ((5, 7), (3, 4)), M=3, D=8, P=3, regime="O",
                                                  seed=1, Lq=9)
    v = value.double().requires_grad_(True)
    lo = loc.double().requires_grad_(True)
    at = attn.double().requires_grad_(True)
    out = ms_deform_attn_core_pytorch(v, sh, lo, at)
    g = torch.randn_like(out)
    out.backward(g)
    ref = oracle.forward_c(value, sh, lsi, loc, attn)
    gv, gl, ga = oracle.backward_c(value, sh, lsi, loc, attn, g)
    assert np.abs(out.detach().numpy() - ref).max() < 1e-12
    assert np.abs(v.grad.numpy() - gv).max() < 1e-12
    assert np.abs(lo.grad.numpy() - gl).max() < 1e-10
    assert np.abs(at.grad.numpy() - ga).max() < 1e-12


def test_synth_shapes_and_reference_points():
    from pctrans_b200 import synth
    sh, lsi, S = synth.level_tensors(synth.CVPPP_512)
    assert S == 5376 and lsi.tolist() == [0, 256, 1280]
    sh, lsi, S = synth.level_tensors(synth.BBBC_520x696)
    assert S == 7481
    ref = synth.encoder_reference_points(((2, 3),))
    assert torch.allclose(ref[0], torch.tensor([0.5 / 3, 0.25])) and ref.shape == (6, 2)
    value, sh, lsi, loc, attn = synth.make_inputs(2, synth.CVPPP_448, M=8, D=16, P=4, seed=0)
    assert value.shape == (2, 4116, 8, 16) and loc.shape == (2, 4116, 8, 3, 4, 2)
    assert torch.allclose(attn.sum((-1, -2)), torch.ones(2, 4116, 8), atol=1e-5)


def test_linear_precision_policy_without_gpu():
    """Host logic of functions/linear_tc.py: which operand precision a tensor-core Linear would
    take (PRECISION, torch.autocast) and that CPU tensors are never eligible (they go to
    torch.nn.functional.linear as in the reference)."""
    from pctrans_b200.functions import linear_tc
    assert linear_tc.PRECISION == "auto" and linear_tc.BF16_WGRAD
    assert not linear_tc.bf16_operands()
    if torch.cuda.is_available():       # without CUDA torch disables "cuda" autocast regions
        with torch.autocast("cuda", dtype=torch.bfloat16):
            assert linear_tc.bf16_operands()
    with torch.autocast("cuda", dtype=torch.float16):
        assert not linear_tc.bf16_operands()
    linear_tc.PRECISION = "bf16"
    try:
        assert linear_tc.bf16_operands()
    finally:
        linear_tc.PRECISION = "auto"
    linear_tc.PRECISION = "fp32"
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            assert not linear_tc.bf16_operands()
    finally:
        linear_tc.PRECISION = "auto"
    lin = torch.nn.Linear(32, 32)
    x = torch.randn(4, 32)
    assert not linear_tc.eligible(x, lin.weight)
    assert torch.equal(linear_tc.linear(x, lin.weight, lin.bias, relu=True), torch.relu(lin(x)))
