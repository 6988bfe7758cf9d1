"""GPU parity of the cell-binned backward (pctrans_b200/csrc/msda_bwd_binned.cuh) through the C ABI.

The binned kernel is the default gradient kernel wherever the shape fits it (round 2, second
form: bilinear-basis sweep in lock-step rounds); these cases pin it against (a) the CPU oracle
(fp64 closed form of cuh:93-164) and (b) the per-tap scatter kernel it replaces
(``PCT_MSDA_FLAG_BWD_PER_TAP``), over the shapes that stress its bookkeeping: several chunks per
(image, head), chunk sizes that do not divide Lq, L*P that does not divide the CTA size, one-pixel
maps, out-of-range samples, Lq != S, bf16 values, every supported head dim.
"""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as oracle
from pctrans_b200 import _cabi, synth

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rel(a, b):
    a = a.detach().double().cpu().numpy()
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-30))


def qc_flag(qc):
    """tuning bits 25..31 of the flags: cap of the chunk size in units of 8 queries"""
    assert qc % 8 == 0 and 0 < qc <= 127 * 8
    return (qc // 8) << 25


def run_case(N, shapes, M, D, P, regime, seed, dtype=torch.float32, Lq=None, flags=0, tol=1e-4,
             nudge=True, jitter_px=1.0):
    value, sh, lsi, loc, attn = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime=regime, seed=seed,
                                                  Lq=Lq, jitter_px=jitter_px)
    if nudge:
        loc = synth.nudge_off_grid(loc, shapes)
    Lq = loc.shape[1]
    g = torch.randn(N, Lq, M * D, generator=torch.Generator().manual_seed(seed + 100))
    if dtype == torch.bfloat16:
        value, g = value.bfloat16().float(), g.bfloat16().float()
    S, L = value.shape[1], len(shapes)
    flags |= _cabi.FLAG_BWD_BINNED
    assert _cabi.backward_plan(N, S, M, D, L, Lq, P, dtype, flags) == 2, "binned kernel not selected"
    assert _cabi.backward_plan(N, S, M, D, L, Lq, P, dtype, _cabi.FLAG_BWD_PER_TAP) == 1
    args = (value.to(DEV, dtype), sh.to(DEV), lsi.to(DEV), loc.to(DEV), attn.to(DEV), g.to(DEV, dtype))
    gv, gl, ga = _cabi.backward(*args, flags=flags)
    pv, pl, pa = _cabi.backward(*args, flags=_cabi.FLAG_BWD_PER_TAP)
    torch.cuda.synchronize()
    rv, rl, ra = oracle.backward_c(value, sh, lsi, loc, attn, g)
    errs = dict(gv=rel(gv, rv), gl=rel(gl, rl), ga=rel(ga, ra),
                gv_vs_tap=rel(gv, pv.double().cpu().numpy()), gl_vs_tap=rel(gl, pl.double().cpu().numpy()),
                ga_vs_tap=rel(ga, pa.double().cpu().numpy()))
    for k, v in errs.items():
        assert v <= tol, (k, errs)
    return errs


@pytest.mark.parametrize("D", [16, 32, 64, 128])
@pytest.mark.parametrize("regime", ["U", "O"])
def test_binned_small_fp32(D, regime):
    run_case(3, ((7, 9), (4, 5), (2, 3)), M=4, D=D, P=4, regime=regime, seed=D + 1, Lq=37)


@pytest.mark.parametrize("D", [16, 32, 64])
def test_binned_small_bf16(D):
    run_case(2, ((6, 11), (3, 6)), M=4, D=D, P=4, regime="O", seed=D, dtype=torch.bfloat16, Lq=53,
             tol=2e-2)


@pytest.mark.parametrize("qc", [8, 24, 64, 200, 0])
def test_binned_chunking(qc):
    """several chunks per (image, head), ragged last chunk, realistic locations (dense cells)"""
    flags = qc_flag(qc) if qc else 0
    run_case(2, ((5, 6), (10, 12), (20, 24)), M=8, D=32, P=4, regime="R", seed=qc + 3, flags=flags)


@pytest.mark.parametrize("LP", [(1, 1), (2, 3), (3, 4), (4, 5), (5, 7), (8, 8)])
def test_binned_sample_counts(LP):
    """L*P from 1 to 64: the sample -> (query, sample) walk of the binning passes"""
    L, P = LP
    shapes = tuple((3 + l, 2 + 2 * l) for l in range(L))
    run_case(2, shapes, M=2, D=32, P=P, regime="O", seed=L * 10 + P, Lq=41)


def test_binned_one_pixel_and_line_maps():
    run_case(2, ((1, 1), (1, 7), (5, 1)), M=2, D=32, P=4, regime="O", seed=5, Lq=64)


def test_binned_on_grid_initial_offsets():
    """mod.py:66-74: at initialisation the offsets are whole pixels, i.e. every sample sits ON a
    cell boundary (lw or lh exactly 0).  No nudging, no jitter: compared with the oracle's literal
    fp32 instantiation, which picks the same cells as the kernels."""
    shapes = synth.CVPPP_512
    value, sh, lsi, loc, attn = synth.make_inputs(1, shapes, M=8, D=32, P=4, regime="R", seed=2,
                                                  jitter_px=0.0)
    g = torch.randn(1, loc.shape[1], 256, generator=torch.Generator().manual_seed(9))
    args = (value.to(DEV), sh.to(DEV), lsi.to(DEV), loc.to(DEV), attn.to(DEV), g.to(DEV))
    gv, gl, ga = _cabi.backward(*args, flags=_cabi.FLAG_BWD_BINNED)
    pv, pl, pa = _cabi.backward(*args, flags=_cabi.FLAG_BWD_PER_TAP)
    rv, rl, ra = oracle.backward_c(value, sh, lsi, loc, attn, g, dtype=np.float32)
    for name, a, p, r in (("gv", gv, pv, rv), ("gl", gl, pl, rl), ("ga", ga, pa, ra)):
        assert rel(a, r) <= 1e-4, name
        assert rel(a, p.double().cpu().numpy()) <= 1e-4, name


def test_binned_cfg3_full_size_vs_per_tap_and_oracle():
    """BASELINE config 3 shape (batch 16 is the bench; 4 images here keep the oracle in seconds)"""
    errs = run_case(4, synth.CVPPP_512, M=8, D=32, P=4, regime="R", seed=0)
    print(errs)


def test_binned_bbbc_d16():
    """non-square, non-power-of-two maps at PCTrans's real head dim (CONVS_DIM 128 -> D = 16)"""
    run_case(2, synth.BBBC_520x696, M=8, D=16, P=4, regime="R", seed=4)


def test_binned_zero_weights_and_nan_free():
    """attention weight 0 still yields grad_attn; nothing of a skipped corner leaks"""
    shapes = ((4, 4), (8, 8))
    value, sh, lsi, loc, attn = synth.make_inputs(1, shapes, M=2, D=32, P=2, regime="O", seed=8, Lq=30)
    attn = attn.clone()
    attn[:, ::2] = 0.0
    loc = synth.nudge_off_grid(loc, shapes)
    g = torch.randn(1, 30, 64, generator=torch.Generator().manual_seed(3))
    gv, gl, ga = _cabi.backward(value.to(DEV), sh.to(DEV), lsi.to(DEV), loc.to(DEV), attn.to(DEV), g.to(DEV),
                                flags=_cabi.FLAG_BWD_BINNED)
    rv, rl, ra = oracle.backward_c(value, sh, lsi, loc, attn, g)
    assert rel(gv, rv) <= 1e-4 and rel(gl, rl) <= 1e-4 and rel(ga, ra) <= 1e-4
    assert torch.isfinite(gv).all() and torch.isfinite(gl).all() and torch.isfinite(ga).all()


def test_binned_back_to_back_launches_share_nothing():
    """the task counter ring: many launches in a row, results identical to a single launch up to
    the atomic ordering of grad_value and bit-identical for grad_loc / grad_attn"""
    value, sh, lsi, loc, attn = synth.make_inputs(2, synth.CVPPP_448, M=8, D=16, P=4, regime="R", seed=6)
    g = torch.randn(2, loc.shape[1], 128, generator=torch.Generator().manual_seed(6))
    args = (value.to(DEV), sh.to(DEV), lsi.to(DEV), loc.to(DEV), attn.to(DEV), g.to(DEV))
    first = _cabi.backward(*args, flags=_cabi.FLAG_BWD_BINNED)
    for _ in range(20):
        last = _cabi.backward(*args, flags=_cabi.FLAG_BWD_BINNED)
    torch.cuda.synchronize()
    assert torch.allclose(first[0], last[0], rtol=0, atol=1e-5 * float(first[0].abs().max()))
    # grad_loc / grad_attn depend on the order of the records inside a cell only through the
    # group that processes them -- each record is reduced on its own, so they are reproducible
    assert torch.equal(first[1], last[1]) and torch.equal(first[2], last[2])


def test_binned_is_the_default_where_it_fits():
    for D in (16, 32, 64, 128):
        for dt in (torch.float32, torch.bfloat16):
            assert _cabi.backward_plan(2, 5376, 8, D, 3, 5376, 4, dt, 0) == 2
    assert _cabi.backward_plan(2, 5376, 8, 32, 3, 5376, 4, torch.float32, _cabi.FLAG_BWD_PER_TAP) == 1
    assert _cabi.backward_plan(2, 5376, 8, 32, 3, 5376, 4, torch.float32, _cabi.FLAG_BWD_BINNED) == 2
    assert _cabi.backward_plan(2, 5376, 8, 30, 3, 5376, 4, torch.float32, _cabi.FLAG_BWD_BINNED) == 0
    # more pixels per image than the record header holds (15 bits): per-tap kernel
    assert _cabi.backward_plan(1, 40000, 8, 32, 1, 100, 4, torch.float32, 0) == 1


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("D,dtype", [(32, torch.float32), (16, torch.float32), (32, torch.bfloat16)])
def test_fused_binned_matches_fused_per_tap(ref_dim, D, dtype):
    """Fused prologue (raw offsets / logits in, d/d(offsets) and the softmax backward out) inside
    the binned kernel, several chunks per (image, head): against the fused per-tap kernel and --
    through tests/test_gpu_fused.py's composed fp64 reference -- against the oracle."""
    from test_gpu_fused import make_case, reference
    shapes, M, P = synth.CVPPP_448, 4, 4
    value, sh, lsi, ref, off, logits, gout, norm = make_case(2, shapes, M, D, P, ref_dim, seed=11 + D)
    if dtype == torch.bfloat16:
        value, gout = value.bfloat16().float(), gout.bfloat16().float()
    args = (value.to(DEV, dtype), sh.to(DEV), lsi.to(DEV), ref.to(DEV), off.to(DEV), logits.to(DEV),
            gout.to(DEV, dtype))
    bv, bo, bl = _cabi.fused_backward(*args, flags=_cabi.FLAG_BWD_BINNED)
    pv, po, pl = _cabi.fused_backward(*args, flags=_cabi.FLAG_BWD_PER_TAP)
    dv, do, dl = _cabi.fused_backward(*args)                      # the default is the binned kernel
    torch.cuda.synchronize()
    assert not torch.equal(bo, po)                                # two different kernels ran
    assert torch.equal(bo, do) and torch.equal(bl, dl)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    _, r_gv, r_goff, r_glog = reference(value, sh, lsi, ref, off, logits, gout, norm, P)
    for name, got, per_tap, want in (("grad_value", bv, pv, r_gv), ("grad_offsets", bo, po, r_goff),
                                     ("grad_logits", bl, pl, r_glog)):
        assert rel(got, want) <= tol, (name, rel(got, want))
        assert rel(got, per_tap.float().cpu().numpy()) <= 2 * tol, name


def test_binned_largest_map_that_fits():
    """Maximum sizes: one 130 x 140 level.  The level shapes live in device memory, so the host sizes
    shared memory from the bound 2S + 2L cells (6 bytes each): S = 18 200 is about the largest map
    that leaves room for a 64-query chunk (the kernel then finds 18 471 real cells and takes ~800
    queries per chunk); beyond that, and beyond 2^31 elements per tensor, the per-tap kernel runs."""
    shapes = ((130, 140),)
    S = 130 * 140
    errs = run_case(1, shapes, 1, 16, 4, "U", 21, Lq=1700)
    assert errs["gv"] <= 1e-4
    assert _cabi.backward_plan(1, S, 1, 16, 1, 100, 4, torch.float32, 0) == 2
    assert _cabi.backward_plan(1, 19000, 1, 16, 1, 100, 4, torch.float32, 0) == 1     # counters do not fit
    assert _cabi.backward_plan(1, 32768, 1, 16, 1, 100, 4, torch.float32, 0) == 1
    assert _cabi.backward_plan(30000, S, 1, 16, 1, 100, 4, torch.float32, 0) == 1     # N S M D >= 2^31: 64-bit offsets


def test_host_shapes_extend_the_binned_kernel_to_large_maps():
    """With a host copy of spatial_shapes (pct_msda_backward_ex2) the exact cell count replaces
    the 2S + 2L bound: a 1024 x 1024 input's pyramid (128^2 + 64^2 + 32^2 = 21 504 pixels, 21 955
    cells) runs on the binned kernel -- same results as the per-tap kernel and the oracle -- and a
    host copy that does not describe the same pixels is an error."""
    shapes = ((32, 32), (64, 64), (128, 128))
    N, M, D, P, Lq = 1, 2, 32, 4, 3000
    value, sh, lsi, loc, attn = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime="U", seed=31, Lq=Lq)
    loc = synth.nudge_off_grid(loc, shapes)
    S = value.shape[1]
    g = torch.randn(N, Lq, M * D, generator=torch.Generator().manual_seed(32))
    assert _cabi.backward_plan(N, S, M, D, 3, Lq, P, torch.float32, 0) == 1
    assert _cabi.backward_plan(N, S, M, D, 3, Lq, P, torch.float32, 0, host_shapes=shapes) == 2
    d_sh = sh.to(DEV)
    args = (value.to(DEV), d_sh, lsi.to(DEV), loc.to(DEV), attn.to(DEV), g.to(DEV))
    hs = _cabi.host_shapes_of(d_sh)
    assert list(hs) == [v for hw in shapes for v in hw]
    bv, bl, ba = _cabi.backward(*args, host_shapes=hs)
    pv, pl, pa = _cabi.backward(*args)                      # no host copy: the per-tap kernel
    torch.cuda.synchronize()
    rv, rl, ra = oracle.backward_c(value, sh, lsi, loc, attn, g)
    assert not torch.equal(bl, pl)
    for got, tap, want in ((bv, pv, rv), (bl, pl, rl), (ba, pa, ra)):
        assert rel(got, want) <= 1e-4 and rel(got, tap.double().cpu().numpy()) <= 1e-4
    import ctypes
    wrong = (ctypes.c_int64 * 6)(32, 32, 64, 64, 128, 127)
    with pytest.raises(_cabi.MSDAError):
        _cabi.backward(*args, host_shapes=wrong)
