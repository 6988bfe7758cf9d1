"""Tensor-core Linear (include/pct_linear.h) against fp64 references.

The bar is the reference's fp32 one (the pixel decoder runs in fp32 with autocast off,
PD/msdeformattn.py:314-320): results must sit at fp32-GEMM accuracy, not TF32 accuracy.  Tolerances
are relative to the largest reference magnitude: 1e-5 forward (BASELINE's fp32 forward bar),
1e-4 for gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _case(rows, k, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    b = torch.randn(n, generator=g).cuda()
    res = torch.randn(rows, n, generator=g).cuda()
    return x, w, b, res


@pytest.mark.parametrize("rows,k,n", [
    (1, 16, 16), (127, 32, 48), (128, 256, 256), (129, 256, 192), (1000, 256, 96),
    (777, 256, 1024), (513, 1024, 256), (300, 128, 288), (5376, 256, 256)])
def test_forward_epilogues_match_fp64(rows, k, n):
    from pctrans_b200 import _cabi
    x, w, b, res = _case(rows, k, n)
    wp = _cabi.linear_pack_weight(w)
    ref = x.double() @ w.double().t()
    assert _rel(_cabi.linear_forward(x, wp, n), ref) < 1e-5
    ref_b = ref + b.double()
    assert _rel(_cabi.linear_forward(x, wp, n, b), ref_b) < 1e-5
    assert _rel(_cabi.linear_forward(x, wp, n, b, relu=True), ref_b.relu()) < 1e-5
    assert _rel(_cabi.linear_forward(x, wp, n, b, res), ref_b + res.double()) < 1e-5
    assert _rel(_cabi.linear_forward(x, wp, n, b, res, relu=True), (ref_b + res.double()).relu()) < 1e-5
    if n <= 256:
        gamma, beta = torch.randn(n).cuda(), torch.randn(n).cuda()
        ln = F.layer_norm(ref_b + res.double(), (n,), gamma.double(), beta.double(), 1e-5)
        assert _rel(_cabi.linear_forward(x, wp, n, b, res, gamma, beta, 1e-5), ln) < 1e-5
        ln0 = F.layer_norm(ref_b, (n,), gamma.double(), beta.double(), 1e-5)
        assert _rel(_cabi.linear_forward(x, wp, n, b, None, gamma, beta, 1e-5), ln0) < 1e-5


@pytest.mark.parametrize("rows,k,n", [
    (1, 16, 16), (100, 32, 48), (1000, 256, 96), (4116, 128, 288), (5376, 256, 1024),
    (3000, 1024, 256), (86016, 256, 256)])
def test_weight_gradient_matches_fp64(rows, k, n):
    from pctrans_b200 import _cabi
    x, _, _, _ = _case(rows, k, n, seed=3)
    gy = torch.randn(rows, n, generator=torch.Generator().manual_seed(4)).cuda()
    gw = _cabi.linear_wgrad(gy, x)
    assert gw.shape == (n, k)
    assert _rel(gw, gy.double().t() @ x.double()) < 1e-5
    # fully overwritten: a second call into a dirty buffer gives the same answer; the bias
    # gradient (column sums of gy) comes out of the same pass
    gw2, gb = _cabi.linear_wgrad(gy, x, with_bias=True)
    assert _rel(gw2, gy.double().t() @ x.double()) < 1e-5
    assert _rel(gb, gy.double().sum(0)) < 1e-5


def test_accuracy_is_fp32_class_not_tf32_class():
    from pctrans_b200 import _cabi
    x, w, b, _ = _case(4096, 256, 256, seed=1)
    ref = x.double() @ w.double().t() + b.double()
    ours = _rel(_cabi.linear_forward(x, _cabi.linear_pack_weight(w), 256, b), ref)
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        tf32 = _rel(F.linear(x, w, b), ref)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert ours < 5e-6 and ours * 20 < tf32


def test_untouched_rows_and_columns_stay_untouched():
    """Row / column tails: nothing outside (rows, n) is written."""
    from pctrans_b200 import _cabi
    rows, k, n = 130, 64, 48
    x, w, b, _ = _case(rows, k, n)
    buf = torch.full((rows + 4, n), 7.0).cuda()
    _cabi.linear_forward(x, _cabi.linear_pack_weight(w), n, b, out=buf[:rows])
    torch.cuda.synchronize()
    assert bool((buf[rows:] == 7.0).all())
    assert _rel(buf[:rows], x.double() @ w.double().t() + b.double()) < 1e-5


def test_bad_arguments_are_error_codes():
    from pctrans_b200 import _cabi
    x = torch.randn(8, 24).cuda()
    with pytest.raises(_cabi.MSDAError):
        _cabi.linear_pack_weight(torch.randn(16, 24).cuda())          # in % 16
    wp = _cabi.linear_pack_weight(torch.randn(512, 32).cuda())
    g = torch.ones(512).cuda()
    with pytest.raises(_cabi.MSDAError) as e:                           # LayerNorm wider than a tile
        _cabi.linear_forward(torch.randn(8, 32).cuda(), wp, 512, None, None, g, g)
    assert e.value.code == _cabi.E_UNSUPPORTED
    wp2 = _cabi.linear_pack_weight(torch.randn(32, 32).cuda())
    flat = torch.randn(8 * 32 + 1).cuda()
    with pytest.raises(_cabi.MSDAError) as e:                           # x not 16-byte aligned
        _cabi.linear_forward(flat[1:].view(8, 32), wp2, 32)
    assert e.value.code == 1004
    del x


def test_autograd_function_matches_fp64_gradients():
    from pctrans_b200.functions import linear_tc
    rows, k, n = 700, 256, 192
    x, w, b, res = _case(rows, k, n, seed=2)
    gy = torch.randn(rows, n).cuda()
    for relu in (False, True):
        leaves = [t.clone().requires_grad_(True) for t in (x, w, b, res)]
        y = linear_tc.linear(leaves[0].view(7, 100, k), leaves[1], leaves[2],
                             residual=leaves[3].view(7, 100, n), relu=relu)
        assert y.shape == (7, 100, n) and isinstance(y.grad_fn, torch.autograd.function.BackwardCFunction)
        y.backward(gy.view(7, 100, n))
        ref_leaves = [t.double().clone().requires_grad_(True) for t in (x, w, b, res)]
        yr = F.linear(ref_leaves[0], ref_leaves[1], ref_leaves[2]) + ref_leaves[3]
        yr = yr.relu() if relu else yr
        yr.backward(gy.double())
        assert _rel(y.reshape(rows, n), yr.detach()) < 1e-5
        for ours, ref in zip(leaves, ref_leaves):
            assert _rel(ours.grad, ref.grad) < 1e-4


def test_fused_layernorm_autograd_matches_fp64_gradients():
    """norm(residual + linear(x)) as one kernel in training: output and all six gradients."""
    from pctrans_b200.functions import linear_tc
    rows, k, n = 900, 1024, 256
    x, w, b, res = _case(rows, k, n, seed=5)
    norm = torch.nn.LayerNorm(n).cuda()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_()
    gy = torch.randn(rows, n).cuda()
    leaves = [t.clone().requires_grad_(True) for t in (x, w, b, res)]
    y = linear_tc.linear_residual_layernorm(leaves[0].view(9, 100, k), leaves[1], leaves[2],
                                            leaves[3].view(9, 100, n), norm)
    assert type(y.grad_fn).__name__ == "LinearResidualLayerNormFunctionBackward"
    y.backward(gy.view(9, 100, n))
    ref_leaves = [t.double().clone().requires_grad_(True) for t in (x, w, b, res)]
    lw, lb = (p.detach().double().clone().requires_grad_(True) for p in (norm.weight, norm.bias))
    yr = F.layer_norm(F.linear(ref_leaves[0], ref_leaves[1], ref_leaves[2]) + ref_leaves[3], (n,),
                      lw, lb, norm.eps)
    yr.backward(gy.double())
    assert _rel(y.detach().reshape(rows, n), yr.detach()) < 1e-5
    for ours, ref in zip(leaves + [norm.weight, norm.bias], ref_leaves + [lw, lb]):
        assert _rel(ours.grad, ref.grad) < 1e-4


def test_cta_pair_kernel_gives_the_same_results():
    """The opt-in cta_group::2 variant (PCT_LINEAR_PAIR=1, read once per process): same shapes and
    epilogues, in a subprocess, against fp64."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys; sys.path.insert(0, %r)
import torch, torch.nn.functional as F
from pctrans_b200 import _cabi
torch.manual_seed(0)
worst = 0.0
for rows, k, n in [(1, 16, 16), (129, 256, 192), (300, 64, 48), (1000, 256, 96), (777, 256, 1024),
                   (513, 1024, 256), (5376, 256, 256), (5000, 128, 288)]:
    x = torch.randn(rows, k).cuda(); w = (torch.randn(n, k) / k ** 0.5).cuda()
    b = torch.randn(n).cuda(); res = torch.randn(rows, n).cuda()
    wp = _cabi.linear_pack_weight(w)
    ref = x.double() @ w.double().t() + b.double()
    outs = [(_cabi.linear_forward(x, wp, n, b, relu=True), ref.relu()),
            (_cabi.linear_forward(x, wp, n, b, res), ref + res.double())]
    if n <= 256:
        g, be = torch.rand(n).cuda() + 0.5, torch.randn(n).cuda()
        outs.append((_cabi.linear_forward(x, wp, n, b, res, g, be, 1e-5),
                     F.layer_norm(ref + res.double(), (n,), g.double(), be.double(), 1e-5)))
    for y, r in outs:
        worst = max(worst, float((y.double() - r).abs().max() / r.abs().max()))
print("PAIR_WORST", worst)
""" % root
    env = dict(os.environ, PCT_LINEAR_PAIR="1")
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    worst = float(proc.stdout.split("PAIR_WORST")[1])
    assert worst < 1e-5, worst


def test_deterministic_mode_gives_bit_identical_weight_gradients():
    """The split-K wgrad combines slices with fp32 reductions; under
    torch.use_deterministic_algorithms the Function must not use it."""
    from pctrans_b200.functions import linear_tc
    x, w, b, _ = _case(20000, 256, 192, seed=6)
    gy = torch.randn(20000, 192).cuda()

    def grads():
        leaves = [t.clone().requires_grad_(True) for t in (x, w, b)]
        linear_tc.linear(leaves[0], leaves[1], leaves[2]).backward(gy)
        return [t.grad.clone() for t in leaves]

    torch.use_deterministic_algorithms(True)
    try:
        first, second = grads(), grads()
    finally:
        torch.use_deterministic_algorithms(False)
    for a_, b_ in zip(first, second):
        assert torch.equal(a_, b_)
    ref_w = gy.double().t() @ x.double()
    assert _rel(first[1], ref_w) < 1e-4


def test_encoder_inference_replays_from_a_cuda_graph():
    """Kernels, packed-weight cache and fused epilogues under CUDA graph capture: a captured
    3-layer encoder forward replays to the eager result (launch-bound inference loops, B200 guide)."""
    from pctrans_b200 import encoder as enc, synth
    torch.manual_seed(1)
    pyramid = synth.CVPPP_448
    model = enc.MSDeformAttnTransformerEncoderOnly(
        d_model=128, nhead=8, num_encoder_layers=3, dim_feedforward=512, dropout=0.0,
        num_feature_levels=len(pyramid)).cuda().eval()
    srcs = [torch.randn(2, 128, h, w).cuda() for h, w in pyramid]
    pos = [enc.PositionEmbeddingSine(64, normalize=True)(s) for s in srcs]
    with torch.no_grad():
        eager = model(srcs, pos)[0].clone()          # also warms the packed-weight cache
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                model(srcs, pos)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = model(srcs, pos)[0]
        for s in srcs:
            s.mul_(0.5)                               # new inputs in the captured buffers
        graph.replay()
        torch.cuda.synchronize()
        expect = model(srcs, pos)[0]
    assert _rel(captured, expect.double()) < 1e-6
    assert not torch.equal(captured, eager)


def test_gated_epilogue_is_relu_backward():
    from pctrans_b200 import _cabi
    x, w, b, _ = _case(1000, 256, 1024, seed=7)
    gate = torch.randn(1000, 1024).cuda()
    y = _cabi.linear_forward(x, _cabi.linear_pack_weight(w), 1024, b, gate=gate)
    ref = (x.double() @ w.double().t() + b.double()) * (gate > 0)
    assert _rel(y, ref) < 1e-5
    assert bool((y[gate <= 0] == 0).all())


def test_ffn_block_autograd_matches_fp64():
    """norm(src + linear2(relu(linear1(src)))) as FFNFunction: output and all seven gradients."""
    from pctrans_b200.functions import linear_tc
    torch.manual_seed(8)
    lin1, lin2, norm = torch.nn.Linear(256, 1024).cuda(), torch.nn.Linear(1024, 256).cuda(), torch.nn.LayerNorm(256).cuda()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_()
    src = torch.randn(3, 700, 256).cuda().requires_grad_(True)
    gy = torch.randn(3, 700, 256).cuda()
    y = linear_tc.ffn_block(src, lin1, lin2, norm)
    assert type(y.grad_fn).__name__ == "FFNFunctionBackward"
    y.backward(gy)
    ours = [src.grad] + [p.grad for p in (lin1.weight, lin1.bias, lin2.weight, lin2.bias, norm.weight, norm.bias)]
    d = [t.detach().double().clone().requires_grad_(True)
         for t in (src, lin1.weight, lin1.bias, lin2.weight, lin2.bias, norm.weight, norm.bias)]
    # ReLU's gradient is discontinuous at 0: a pre-activation within fp32 rounding of 0 (a handful
    # of the 2.1 M here) may get the other sign in fp64, so the reference uses OUR activation mask
    with torch.no_grad():
        mask = (linear_tc.linear(src.detach(), lin1.weight, lin1.bias, relu=True) > 0).double()
    yr = F.layer_norm(d[0] + F.linear(F.linear(d[0], d[1], d[2]) * mask, d[3], d[4]), (256,), d[5], d[6], norm.eps)
    yr.backward(gy.double())
    assert _rel(y.detach(), yr.detach()) < 1e-5
    for a_, r_ in zip(ours, d):
        assert _rel(a_, r_.grad) < 1e-4


def test_packed_weight_cache_follows_in_place_updates():
    from pctrans_b200.functions import linear_tc
    lin = torch.nn.Linear(64, 32).cuda()
    x = torch.randn(50, 64).cuda()
    with torch.no_grad():
        y0 = linear_tc.linear(x, lin.weight, lin.bias)
        p0 = linear_tc.packed_weight(lin.weight)
        assert linear_tc.packed_weight(lin.weight) is p0            # cached
        lin.weight.mul_(2.0)                                         # optimizer-style in-place step
        y1 = linear_tc.linear(x, lin.weight, lin.bias)
    assert linear_tc.packed_weight(lin.weight) is not p0
    assert _rel(y1 - lin.bias, 2.0 * (y0 - lin.bias).double()) < 1e-5


def test_ineligible_inputs_take_the_torch_linear():
    """bf16 / autocast / odd sizes are not this kernel's job: same result as F.linear."""
    from pctrans_b200.functions import linear_tc
    lin = torch.nn.Linear(20, 12).cuda()
    x = torch.randn(9, 20).cuda()
    assert not linear_tc.eligible(x, lin.weight)
    assert torch.equal(linear_tc.linear(x, lin.weight, lin.bias), lin(x))
    lin2 = torch.nn.Linear(32, 32).cuda()
    x2 = torch.randn(9, 32).cuda()
    with torch.autocast("cuda", dtype=torch.float16):                # fp16 autocast: torch's Linear
        assert not linear_tc.eligible(x2, lin2.weight)
        assert linear_tc.linear(x2, lin2.weight, lin2.bias).dtype == torch.float16
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert linear_tc.eligible(x2, lin2.weight)                   # bf16 operands, own kernel
        assert not linear_tc.eligible(x2.bfloat16(), lin2.weight)    # bf16 tensors: torch's Linear
        lin3 = torch.nn.Linear(48, 32).cuda()                        # K % 32 != 0: torch's Linear
        assert not linear_tc.eligible(torch.randn(9, 48).cuda(), lin3.weight)
        linear_tc.PRECISION = "fp32"                                 # pinned to 3xTF32: not under autocast
        try:
            assert not linear_tc.eligible(x2, lin2.weight)
            assert linear_tc.linear(x2, lin2.weight, lin2.bias).dtype == torch.bfloat16
        finally:
            linear_tc.PRECISION = "auto"


@pytest.mark.parametrize("mode", ["infer", "train"])
def test_encoder_with_tensor_core_linears_matches_torch_linears(mode):
    """The 6-layer encoder with the tensor-core projections / FFN (and, without autograd, the fused
    residual + LayerNorm epilogues) against the same weights through torch's fp32 Linears."""
    from pctrans_b200 import encoder as enc, synth
    from pctrans_b200.functions import linear_tc
    torch.manual_seed(0)
    pyramid = synth.CVPPP_448
    model = enc.MSDeformAttnTransformerEncoderOnly(
        d_model=128, nhead=8, num_encoder_layers=3, dim_feedforward=1024, dropout=0.0,
        num_feature_levels=len(pyramid)).cuda()
    with torch.no_grad():                     # generic offsets, not the on-grid initial ones
        for m in model.modules():
            if isinstance(m, enc.MSDeformAttn):
                m.sampling_offsets.weight.normal_(0, 0.02)
                m.attention_weights.weight.normal_(0, 0.1)
    srcs = [torch.randn(2, 128, h, w).cuda() for h, w in pyramid]
    pos = [enc.PositionEmbeddingSine(64, normalize=True)(s) for s in srcs]
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    outs, grads = [], []
    try:
        for enabled in (True, False):
            linear_tc.ENABLED = enabled
            model.zero_grad(set_to_none=True)
            if mode == "infer":
                with torch.no_grad():
                    outs.append(model(srcs, pos)[0])
            else:
                out = model(srcs, pos)[0]
                out.square().mean().backward()
                outs.append(out.detach())
                grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
    finally:
        linear_tc.ENABLED = True
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
    assert _rel(outs[0], outs[1].double()) < 1e-5
    if mode == "train":
        # relative L2 per parameter.  Same bar as the whole-model gradient test in
        # test_encoder.py: the 1e-6 perturbation of the offsets moves a few samples across pixel
        # boundaries, where the location gradient is discontinuous, so the offset / attention
        # Linears (gradient norms ~1e-6 of the largest) differ by ~1e-2 between ANY two fp32
        # evaluation orders (measured 1.4e-2 worst; the Linears downstream of them up to 5e-3).
        # The per-GEMM gradients are pinned tightly above (1e-4 / 1e-5 against fp64).
        worst = {}
        for k in grads[0]:
            den = float(grads[1][k].double().norm().clamp_min(1e-30))
            worst[k] = float((grads[0][k].double() - grads[1][k].double()).norm()) / den
        bad = {k: v for k, v in worst.items() if v > 3e-2}
        assert not bad, sorted(bad.items(), key=lambda t: -t[1])[:5]


def test_row_mask_epilogue_is_masked_fill_with_gradients():
    """PCT_LINEAR_ROWMASK: value_proj + masked_fill(padding_mask) (mod.py:98-101) as one kernel --
    output rows exactly zero, gradients of x / W / b those of the masked composition (fp64)."""
    from pctrans_b200.functions import linear_tc
    x, w, b, _ = _case(2 * 517, 256, 256, seed=11)
    x = x.view(2, 517, 256)
    mask = torch.zeros(2, 517, dtype=torch.bool, device="cuda")
    mask[0, -40:] = True
    mask[1, 100:130] = True
    xg, wg, bg = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = linear_tc.linear(xg, wg, bg, row_mask=mask)
    assert bool((y[mask] == 0).all()) and bool((y[~mask] != 0).any())
    gy = torch.randn_like(y)
    y.backward(gy)
    xd, wd, bd = (t.double().detach().requires_grad_(True) for t in (x, w, b))
    yd = torch.nn.functional.linear(xd, wd, bd).masked_fill(mask[..., None], 0.0)
    yd.backward(gy.double())
    assert _rel(y, yd) < 1e-5
    assert _rel(xg.grad, xd.grad) < 1e-5 and _rel(wg.grad, wd.grad) < 1e-5 and _rel(bg.grad, bd.grad) < 1e-5
    assert bool((xg.grad[mask] == 0).all())
    # an unaligned mask pointer is fine (bytes), the flag cannot be combined with a residual
    from pctrans_b200 import _cabi
    x2 = x.view(-1, 256)
    m8 = torch.zeros(x2.shape[0] + 1, dtype=torch.uint8, device="cuda")[1:]
    m8[5] = 1
    y2 = _cabi.linear_forward(x2, _cabi.linear_pack_weight(w), 256, b, row_mask=m8)
    assert bool((y2[5] == 0).all()) and bool((y2[6] != 0).any())
    with pytest.raises(ValueError):
        _cabi.linear_forward(x2, _cabi.linear_pack_weight(w), 256, b, residual=y2, row_mask=m8)


@pytest.mark.parametrize("rows,n", [(1000, 256), (86016, 256), (517, 128), (33, 512), (1, 384)])
def test_layernorm_backward_matches_fp64(rows, n):
    """pct_linear_layernorm_backward (one pass: grad of the pre-norm row + column sums for gamma / beta)
    against torch's LayerNorm differentiated in fp64."""
    from pctrans_b200 import _cabi
    g = torch.Generator().manual_seed(rows + n)
    x = (torch.randn(rows, n, generator=g) * 2 + 0.5).cuda()
    gamma = (torch.randn(n, generator=g) * 0.5 + 1).cuda()
    beta = torch.randn(n, generator=g).cuda()
    gy = torch.randn(rows, n, generator=g).cuda()
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    F.layer_norm(xd, (n,), gd, bd, 1e-5).backward(gy.double())
    mean = x.double().mean(1, keepdim=True)
    rstd = (x.double().var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    gp, gg, gb = _cabi.layernorm_backward(gy, x, mean.float().contiguous(), rstd.float().contiguous(), gamma)
    assert _rel(gp, xd.grad) < 2e-5 and _rel(gg, gd.grad) < 2e-5 and _rel(gb, bd.grad) < 2e-5
    assert not _cabi.layernorm_backward_supported(64) and not _cabi.layernorm_backward_supported(1024)


# ---- bf16 operand precision (PCT_LINEAR_BF16) -----------------------------------------------------
# The kernel rounds x and W to bf16 (nearest even) and accumulates exact products in fp32, so against
# an fp64 GEMM of the ROUNDED operands it must be as tight as the 3xTF32 kernel is against the
# unrounded ones (1e-5); against the unrounded fp64 result it sits at bf16 accuracy (BASELINE's
# 2e-2 bar for bf16).

def _r16(t):
    return t.bfloat16().double()


@pytest.mark.parametrize("rows,k,n", [
    (1, 32, 16), (127, 32, 48), (128, 256, 256), (129, 256, 192), (1000, 256, 96),
    (777, 256, 1024), (513, 1024, 256), (300, 128, 288), (5376, 256, 256), (86016, 256, 256)])
def test_bf16_operand_forward_epilogues_match_fp64_of_rounded_operands(rows, k, n):
    from pctrans_b200 import _cabi
    x, w, b, res = _case(rows, k, n, seed=11)
    wp = _cabi.linear_pack_weight(w, bf16=True)
    fwd = lambda *a, **kw: _cabi.linear_forward(x, wp, n, *a, bf16=True, **kw)
    ref = _r16(x) @ _r16(w).t()
    assert _rel(fwd(), ref) < 1e-5
    assert _rel(fwd(), x.double() @ w.double().t()) < 2e-2
    ref_b = ref + b.double()
    assert _rel(fwd(b), ref_b) < 1e-5
    assert _rel(fwd(b, relu=True), ref_b.relu()) < 1e-5
    assert _rel(fwd(b, res), ref_b + res.double()) < 1e-5
    gate = torch.randn(rows, n, generator=torch.Generator().manual_seed(5)).cuda()
    assert _rel(fwd(b, gate=gate), ref_b * (gate > 0)) < 1e-5
    mask = (torch.rand(rows, generator=torch.Generator().manual_seed(6)) < 0.3).cuda()
    assert _rel(fwd(b, row_mask=mask), ref_b.masked_fill(mask[:, None], 0.0)) < 1e-5
    if n <= 256:
        gamma, beta = torch.randn(n).cuda(), torch.randn(n).cuda()
        ln = F.layer_norm(ref_b + res.double(), (n,), gamma.double(), beta.double(), 1e-5)
        y, pre, mean, rstd = fwd(b, res, gamma, beta, 1e-5, save_for_backward=True)
        assert _rel(y, ln) < 1e-5
        assert _rel(pre, ref_b + res.double()) < 1e-5
        assert _rel(mean.view(-1), (ref_b + res.double()).mean(1)) < 1e-5


@pytest.mark.parametrize("rows,k,n", [
    (1, 32, 16), (100, 32, 48), (1000, 256, 96), (4116, 128, 288), (5376, 256, 1024),
    (3000, 1024, 256), (86016, 256, 256)])
def test_bf16_operand_weight_gradient_matches_fp64_of_rounded_operands(rows, k, n):
    from pctrans_b200 import _cabi
    x, _, _, _ = _case(rows, k, n, seed=3)
    gy = torch.randn(rows, n, generator=torch.Generator().manual_seed(4)).cuda()
    ref = _r16(gy).t() @ _r16(x)
    gw = _cabi.linear_wgrad(gy, x, bf16=True)
    assert gw.shape == (n, k)
    assert _rel(gw, ref) < 1e-5
    assert _rel(gw, gy.double().t() @ x.double()) < 2e-2
    gw2, gb = _cabi.linear_wgrad(gy, x, with_bias=True, bf16=True)
    assert _rel(gw2, ref) < 1e-5
    assert _rel(gb, gy.double().sum(0)) < 1e-5           # the bias gradient sums the unrounded grad_y


def test_bf16_operand_bad_arguments_are_error_codes():
    from pctrans_b200 import _cabi
    x, w, _, _ = _case(64, 48, 32)
    with pytest.raises(_cabi.MSDAError):
        _cabi.linear_pack_weight(w, bf16=True)            # in_features % 32 != 0
    wp = _cabi.linear_pack_weight(w)
    with pytest.raises(_cabi.MSDAError):
        _cabi.linear_forward(x, wp, 32, bf16=True)


def test_bf16_operands_under_autocast_with_gradients():
    """Under bf16 autocast the Functions run the bf16-operand kernels forward and for the input
    and weight gradients (exact against fp64 GEMMs of the rounded operands) and remember the mode
    for a backward called outside the autocast region; BF16_WGRAD = False keeps the weight gradient
    on the 3xTF32 kernel."""
    from pctrans_b200 import _cabi
    from pctrans_b200.functions import linear_tc
    x, w, b, res = _case(3000, 256, 192, seed=21)
    x.requires_grad_(True)
    w.requires_grad_(True)
    b.requires_grad_(True)
    n0 = _cabi.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert linear_tc.bf16_operands()
        y = linear_tc.linear(x, w, b, relu=True)
    assert y.dtype == torch.float32 and _cabi.launch_count() > n0
    ref = (_r16(x) @ _r16(w).t() + b.double()).relu()
    assert _rel(y, ref) < 1e-5
    gy = torch.randn_like(y)
    y.backward(gy)                                        # outside the autocast region
    g = gy.double() * (ref > 0)
    assert _rel(x.grad, _r16(g.float()) @ _r16(w)) < 1e-5
    assert _rel(w.grad, _r16(g.float()).t() @ _r16(x)) < 1e-5    # bf16-operand weight gradient
    assert _rel(w.grad, g.t() @ x.double()) < 2e-2
    assert _rel(b.grad, g.sum(0)) < 1e-5
    linear_tc.BF16_WGRAD = False
    try:
        w.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            linear_tc.linear(x, w, b, relu=True).backward(gy)
        assert _rel(w.grad, g.t() @ x.double()) < 1e-5
    finally:
        linear_tc.BF16_WGRAD = True
    # the same call outside autocast is the 3xTF32 kernel
    y32 = linear_tc.linear(x.detach(), w.detach(), b.detach(), relu=True)
    assert _rel(y32, (x.double() @ w.double().t() + b.double()).relu()) < 1e-5


@pytest.mark.parametrize("mode", ["infer", "train"])
def test_encoder_under_bf16_autocast_runs_own_kernels_and_tracks_fp32(mode):
    """Six-layer-style encoder under torch.autocast(bfloat16): every Linear is the bf16-operand
    tensor-core kernel (no cuBLAS: the launch counter advances by the same number of own kernels as
    in the fp32 run), the output tracks the fp32 run within BASELINE's bf16 bar (2e-2)."""
    from pctrans_b200 import _cabi, encoder as enc, synth
    torch.manual_seed(0)
    pyramid = synth.CVPPP_448
    model = enc.MSDeformAttnTransformerEncoderOnly(
        d_model=256, nhead=8, num_encoder_layers=3, dim_feedforward=1024, dropout=0.0,
        num_feature_levels=len(pyramid)).cuda()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, enc.MSDeformAttn):
                m.sampling_offsets.weight.normal_(0, 0.02)
                m.attention_weights.weight.normal_(0, 0.1)
    srcs = [torch.randn(2, 256, h, w).cuda() for h, w in pyramid]
    pos = [enc.PositionEmbeddingSine(128, normalize=True)(s) for s in srcs]

    def run(amp):
        model.zero_grad(set_to_none=True)
        n0 = _cabi.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            if mode == "infer":
                with torch.no_grad():
                    out = model(srcs, pos)[0]
            else:
                out = model(srcs, pos)[0]
        grads = None
        if mode == "train":
            out.square().mean().backward()
            grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        return out.detach(), grads, _cabi.launch_count() - n0

    run(True)                                             # packs the bf16 weights
    o16, g16, n16 = run(True)
    run(False)
    o32, g32, n32 = run(False)
    assert o16.dtype == torch.float32
    assert n16 == n32, (n16, n32)                         # same own kernels, none replaced by cuBLAS
    assert _rel(o16, o32.double()) < 2e-2
    if mode == "train":
        for k in ("encoder.layers.0.linear1.weight", "encoder.layers.2.self_attn.value_proj.weight",
                  "encoder.layers.1.self_attn.output_proj.weight", "encoder.layers.2.norm2.weight"):
            den = float(g32[k].double().norm())
            assert float((g16[k].double() - g32[k].double()).norm()) / den < 5e-2, k


def test_bf16_tensors_in_global_memory_change_no_result():
    """PCT_LINEAR_X_BF16 / Y_BF16 / RES_BF16 / GY_BF16: a tensor that is bf16 in memory gives exactly
    what the fp32 tensor holding the same (rounded) values gives; a bf16 result is the fp32 result
    rounded once."""
    from pctrans_b200 import _cabi
    rows, k, n = 3000, 1024, 256
    x, w, b, res = _case(rows, k, n, seed=31)
    wp = _cabi.linear_pack_weight(w, bf16=True)
    x16, res16 = x.bfloat16(), res.bfloat16()
    y_ref = _cabi.linear_forward(x16.float(), wp, n, b, res16.float(), relu=True, bf16=True)
    y = _cabi.linear_forward(x16, wp, n, b, res16, relu=True, bf16=True)
    assert y.dtype == torch.float32 and torch.equal(y, y_ref)
    y_low = _cabi.linear_forward(x16, wp, n, b, res16, relu=True, bf16=True, out_dtype=torch.bfloat16)
    assert y_low.dtype == torch.bfloat16 and torch.equal(y_low, y_ref.bfloat16())
    gate = torch.randn(rows, n, generator=torch.Generator().manual_seed(7)).cuda()
    g_ref = _cabi.linear_forward(x, wp, n, b, gate=gate, bf16=True)
    assert torch.equal(_cabi.linear_forward(x, wp, n, b, gate=gate.bfloat16(), bf16=True), g_ref)
    # LayerNorm epilogue with a bf16 x: statistics and the saved input stay fp32
    gamma, beta = torch.randn(n).cuda(), torch.randn(n).cuda()
    a1 = _cabi.linear_forward(x16, wp, n, b, res, gamma, beta, 1e-5, save_for_backward=True, bf16=True)
    a0 = _cabi.linear_forward(x16.float(), wp, n, b, res, gamma, beta, 1e-5, save_for_backward=True, bf16=True)
    assert all(t1.dtype == torch.float32 and torch.equal(t1, t0) for t1, t0 in zip(a1, a0))
    # weight gradient: either operand (or both) bf16 in memory
    gy = torch.randn(rows, n, generator=torch.Generator().manual_seed(8)).cuda()
    ref = _r16(gy).t() @ _r16(x)
    for gy_in, x_in in ((gy.bfloat16(), x), (gy, x16), (gy.bfloat16(), x16)):
        gw, gb = _cabi.linear_wgrad(gy_in, x_in, with_bias=True, bf16=True)
        assert gw.dtype == torch.float32 and _rel(gw, ref) < 1e-5
        assert _rel(gb, gy_in.double().sum(0)) < 1e-5
    # bf16 tensors without the bf16 kernel are refused
    with pytest.raises(TypeError):
        _cabi.linear_forward(x16, _cabi.linear_pack_weight(w), n)
    with pytest.raises(TypeError):
        _cabi.linear_wgrad(gy.bfloat16(), x)


def test_ffn_block_bf16_hidden_activation_changes_only_the_bias_gradient():
    """Under bf16 autocast the FFN block stores its hidden activation and that activation's
    gradient in bf16 (BF16_HIDDEN): every GEMM rounds those operands anyway, so the output is
    bit-identical to the fp32-storage run and the gradients agree to fp32-reduction noise; only the
    first bias gradient sums rounded values."""
    from pctrans_b200.functions import linear_tc
    torch.manual_seed(5)
    lin1, lin2 = torch.nn.Linear(256, 1024).cuda(), torch.nn.Linear(1024, 256).cuda()
    norm = torch.nn.LayerNorm(256).cuda()
    src = torch.randn(2, 2500, 256, device="cuda", requires_grad=True)
    gy = torch.randn(2, 2500, 256, device="cuda")
    params = [src, lin1.weight, lin1.bias, lin2.weight, lin2.bias, norm.weight, norm.bias]

    def run(hidden_bf16):
        linear_tc.BF16_HIDDEN = hidden_bf16
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = linear_tc.ffn_block(src, lin1, lin2, norm)
            with torch.no_grad():
                y_inf = linear_tc.ffn_block(src.detach(), lin1, lin2, norm)
        y.backward(gy)
        return y.detach(), y_inf, [p.grad.clone() for p in params]

    try:
        y1, yi1, g1 = run(True)
        y0, yi0, g0 = run(False)
    finally:
        linear_tc.BF16_HIDDEN = True
    assert torch.equal(y1, y0) and torch.equal(yi1, yi0) and torch.equal(yi1, y1)
    for a, b_, name in zip(g1, g0, ("src", "w1", "b1", "w2", "b2", "ln_w", "ln_b")):
        assert _rel(a, b_.double()) < (5e-3 if name == "b1" else 1e-5), name
    # and both track the fp32 block within the bf16 bar
    for p in params:
        p.grad = None
    y32 = linear_tc.ffn_block(src, lin1, lin2, norm)
    y32.backward(gy)
    assert _rel(y1, y32.detach().double()) < 2e-2
    assert _rel(g1[0], src.grad.double()) < 2e-2


@pytest.mark.parametrize("rows,k,n", [(1, 32, 16), (127, 64, 48), (129, 256, 192), (1000, 1024, 96),
                                      (4116, 128, 288), (333, 1024, 1024)])
def test_bf16_tensors_in_memory_on_ragged_shapes(rows, k, n):
    """Partial row tiles, column tiles narrower than 32 and 2-tile outputs through the
    bf16-in-memory instantiations: every combination of bf16 x / residual / gate / result equals
    the all-fp32-tensor call on the same (rounded) values."""
    from pctrans_b200 import _cabi
    x, w, b, res = _case(rows, k, n, seed=41)
    wp = _cabi.linear_pack_weight(w, bf16=True)
    x16, res16 = x.bfloat16(), res.bfloat16()
    gate16 = (res - 0.3).bfloat16()
    for xin in (x16.float(), x16):
        for rin in (res16.float(), res16):
            ref = _cabi.linear_forward(x16.float(), wp, n, b, res16.float(), bf16=True)
            assert torch.equal(_cabi.linear_forward(xin, wp, n, b, rin, bf16=True), ref)
            low = _cabi.linear_forward(xin, wp, n, b, rin, bf16=True, out_dtype=torch.bfloat16)
            assert torch.equal(low, ref.bfloat16())
        gref = _cabi.linear_forward(x16.float(), wp, n, b, gate=gate16.float(), relu=True, bf16=True)
        for gin in (gate16.float(), gate16):
            assert torch.equal(_cabi.linear_forward(xin, wp, n, b, gate=gin, relu=True, bf16=True), gref)
            low = _cabi.linear_forward(xin, wp, n, b, gate=gin, relu=True, bf16=True, out_dtype=torch.bfloat16)
            assert torch.equal(low, gref.bfloat16())
    gy = torch.randn(rows, n, generator=torch.Generator().manual_seed(9)).cuda()
    gw = _cabi.linear_wgrad(gy.bfloat16(), x16, with_bias=True, bf16=True)
    assert _rel(gw[0], _r16(gy).t() @ _r16(x)) < 1e-5 and _rel(gw[1], _r16(gy).sum(0)) < 1e-5


def test_bf16_memory_flags_need_the_bf16_kernel_error_codes():
    """C ABI: PCT_LINEAR_X_BF16 / Y_BF16 / RES_BF16 / GY_BF16 without PCT_LINEAR_BF16, or a bf16
    residual flag with nothing read through the residual slot, are PCT_MSDA_E_UNSUPPORTED."""
    from pctrans_b200 import _cabi
    lib = _cabi.load()
    x, w, b, res = _case(64, 32, 32)
    wp = _cabi.linear_pack_weight(w, bf16=True)
    y = torch.empty(64, 32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def fwd(flags, residual=None):
        return lib.pct_linear_forward(x.data_ptr(), wp.data_ptr(), b.data_ptr(),
                                      residual.data_ptr() if residual is not None else None, None, None,
                                      1e-5, y.data_ptr(), 64, 32, 32, flags, st)
    BF, X16, Y16, R16, G16 = (_cabi.LINEAR_BF16, _cabi.LINEAR_X_BF16, _cabi.LINEAR_Y_BF16,
                              _cabi.LINEAR_RES_BF16, _cabi.LINEAR_GY_BF16)
    assert fwd(BF) == 0
    bad = fwd(X16)
    assert bad != 0 and fwd(Y16) == bad and fwd(R16 | _cabi.LINEAR_RESIDUAL, res) == bad
    assert fwd(BF | R16) == bad                       # no residual / gate to read
    assert fwd(BF | G16) == bad                       # a weight-gradient flag
    gw = torch.empty(32, 32, device="cuda")
    wg = lambda flags: lib.pct_linear_wgrad_ex(res.data_ptr(), x.data_ptr(), gw.data_ptr(), None, 64, 32, 32, flags, st)
    assert wg(0) == 0 and wg(BF) == 0 and wg(X16) == bad and wg(G16) == bad and wg(BF | Y16) == bad
    torch.cuda.synchronize()
    assert lib.pct_msda_error_string(bad).decode()


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("rows,k,na,nb", [(1000, 256, 192, 96), (4116, 128, 192, 96), (129, 64, 32, 16),
                                          (5376, 256, 256, 64)])
def test_pair_of_linears_on_one_input_as_one_gemm(rows, k, na, nb, bf16):
    """pct_linear_forward_split / pct_linear_wgrad_split: sampling_offsets and attention_weights
    (mod.py:102-104) as one GEMM with two dense outputs, one input-gradient GEMM over the pair of
    output gradients and one weight-gradient GEMM -- against the two separate Linears."""
    from pctrans_b200 import _cabi
    g = torch.Generator().manual_seed(51)
    x = torch.randn(rows, k, generator=g).cuda()
    wa, wb = (torch.randn(na, k, generator=g) / k ** 0.5).cuda(), (torch.randn(nb, k, generator=g) / k ** 0.5).cuda()
    ba, bb = torch.randn(na, generator=g).cuda(), torch.randn(nb, generator=g).cuda()
    wcat = torch.cat((wa, wb), 0)
    r = _r16 if bf16 else (lambda t: t.double())
    ya, yb = _cabi.linear_forward_pair(x, _cabi.linear_pack_weight(wcat, bf16=bf16), na, nb,
                                       torch.cat((ba, bb)), bf16=bf16)
    assert _rel(ya, r(x) @ r(wa).t() + ba.double()) < 1e-5
    assert _rel(yb, r(x) @ r(wb).t() + bb.double()) < 1e-5
    # each output column is the same dot product as in the separate Linear: bit-identical
    assert torch.equal(ya, _cabi.linear_forward(x, _cabi.linear_pack_weight(wa, bf16=bf16), na, ba, bf16=bf16))
    assert torch.equal(yb, _cabi.linear_forward(x, _cabi.linear_pack_weight(wb, bf16=bf16), nb, bb, bf16=bf16))
    ga = torch.randn(rows, na, generator=g).cuda()
    gb_ = torch.randn(rows, nb, generator=g).cuda()
    res = torch.randn(rows, k, generator=g).cuda()
    gx_ref = r(ga) @ r(wa) + r(gb_) @ r(wb)
    kstep = 32 if bf16 else 16                 # the pair's widths are the K of the input gradient
    if na % kstep == 0 and (na + nb) % kstep == 0:
        wt = _cabi.linear_pack_weight(wcat.t().contiguous(), bf16=bf16)
        assert _rel(_cabi.linear_forward_cat(ga, gb_, wt, k, bf16=bf16), gx_ref) < 1e-5
        assert _rel(_cabi.linear_forward_cat(ga, gb_, wt, k, residual=res, bf16=bf16), gx_ref + res.double()) < 1e-5
    gw, gbias = _cabi.linear_wgrad_pair(ga, gb_, x, bf16=bf16)
    assert _rel(gw[:na], r(ga).t() @ r(x)) < 1e-5 and _rel(gw[na:], r(gb_).t() @ r(x)) < 1e-5
    assert _rel(gbias, torch.cat((ga, gb_), 1).double().sum(0)) < 1e-5


@pytest.mark.parametrize("amp", [False, True])
def test_linear_pair_function_matches_two_linears(amp):
    """linear_pair (one GEMM forward, one input-gradient GEMM, one weight-gradient GEMM) against the
    two separate tensor-core Linears: identical outputs, gradients within fp32 reduction noise; the
    MSDeformAttn module takes it for sampling_offsets / attention_weights."""
    from pctrans_b200 import _cabi
    from pctrans_b200.functions import linear_tc
    torch.manual_seed(3)
    la, lb = torch.nn.Linear(256, 192).cuda(), torch.nn.Linear(256, 96).cuda()
    x = torch.randn(2, 1500, 256, device="cuda", requires_grad=True)
    ga, gb = torch.randn(2, 1500, 192, device="cuda"), torch.randn(2, 1500, 96, device="cuda")
    params = [x, la.weight, la.bias, lb.weight, lb.bias]

    def run(pair):
        for p in params:
            p.grad = None
        n0 = _cabi.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            assert linear_tc.pair_eligible(x, la, lb)
            if pair:
                ya, yb = linear_tc.linear_pair(x, la, lb)
            else:
                ya, yb = linear_tc.linear(x, la.weight, la.bias), linear_tc.linear(x, lb.weight, lb.bias)
        torch.autograd.backward((ya, yb), (ga, gb))
        return ya.detach(), yb.detach(), [p.grad.clone() for p in params], _cabi.launch_count() - n0

    run(True)
    ya1, yb1, g1, n1 = run(True)
    run(False)
    ya0, yb0, g0, n0 = run(False)
    assert torch.equal(ya1, ya0) and torch.equal(yb1, yb0)
    assert n1 < n0                                     # fewer kernels
    for a, b_, name in zip(g1, g0, ("x", "w_a", "b_a", "w_b", "b_b")):
        assert _rel(a, b_.double()) < 1e-5, name
    # in-place weight updates are seen by the pair's cached operand
    with torch.no_grad():
        la.weight.mul_(2.0)
        lb.bias.add_(1.0)
    ya2, yb2 = linear_tc.linear_pair(x.detach(), la, lb)
    assert _rel(ya2 - la.bias, 2.0 * (ya0 - la.bias).double()) < (2e-2 if amp else 1e-5)
    assert _rel(yb2, (yb0 + 1.0).double()) < (2e-2 if amp else 1e-5)

