"""GPU: the fused-prologue kernels (softmax over L*P + location arithmetic inside the op,
SURVEY.md 8f rank 1) against the oracle composed with an fp64 softmax / location prologue."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as oracle
from pctrans_b200 import synth
from pctrans_b200.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
from pctrans_b200.modules import MSDeformAttn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-30))


def make_case(N, shapes, M, D, P, ref_dim, seed, Lq=None):
    """Raw offsets / logits / reference points whose locations stay off the pixel grid."""
    value, sh, lsi, loc, _ = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime="O", seed=seed, Lq=Lq)
    g = torch.Generator().manual_seed(seed + 100)
    L = len(shapes)
    Lq = loc.shape[1]
    loc = synth.nudge_off_grid(loc, shapes, eps=2e-3)
    logits = torch.randn(N, Lq, M, L * P, generator=g) * 2.0
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)          # (L,2) = (W,H)
    if ref_dim == 2:
        ref = torch.rand(N, Lq, L, 2, generator=g)
        off = (loc - ref[:, :, None, :, None, :]) * norm[None, None, None, :, None, :]
    else:
        ref = torch.cat([torch.rand(N, Lq, L, 2, generator=g),
                         torch.rand(N, Lq, L, 2, generator=g) * 0.5 + 0.1], -1)
        off = (loc - ref[:, :, None, :, None, :2]) * P / (ref[:, :, None, :, None, 2:] * 0.5)
    gout = torch.randn(N, Lq, M * D, generator=g)
    return value, sh, lsi, ref.contiguous(), off.contiguous(), logits, gout, norm


def reference(value, sh, lsi, ref, off, logits, gout, norm, P):
    """fp64 prologue (mod.py:104-112) + C oracle + manual chain rule."""
    N, Lq, M, L = off.shape[:4]
    off64, ref64 = off.double(), ref.double()
    if ref.shape[-1] == 2:
        scale = 1.0 / norm.double()[None, None, None, :, None, :]
        loc = ref64[:, :, None, :, None, :] + off64 * scale
        dloc = scale.expand_as(off64)
    else:
        dloc = (ref64[:, :, None, :, None, 2:] * 0.5 / P).expand_as(off64)
        loc = ref64[:, :, None, :, None, :2] + off64 * dloc
    attn = torch.softmax(logits.double(), -1).view(N, Lq, M, L, P)
    out = oracle.forward_c(value, sh, lsi, loc, attn)
    gv, gl, ga = oracle.backward_c(value, sh, lsi, loc, attn, gout)
    goff = torch.from_numpy(gl) * dloc
    ga_t = torch.from_numpy(ga).view(N, Lq, M, L * P)
    a = attn.view(N, Lq, M, L * P)
    glog = a * (ga_t - (a * ga_t).sum(-1, keepdim=True))
    return out, gv, goff.numpy(), glog.numpy()


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("D,dtype", [(32, torch.float32), (16, torch.float32), (4, torch.float32),
                                     (128, torch.float32), (32, torch.bfloat16), (64, torch.bfloat16)])
def test_fused_matches_composed_oracle(ref_dim, D, dtype):
    shapes, M, P = ((9, 7), (5, 4), (3, 2)), 4, 4
    value, sh, lsi, ref, off, logits, gout, norm = make_case(3, shapes, M, D, P, ref_dim, seed=D + ref_dim, Lq=45)
    if dtype == torch.bfloat16:
        value, gout = value.bfloat16().float(), gout.bfloat16().float()
    v = value.to(DEV, dtype).requires_grad_(True)
    o = off.to(DEV).requires_grad_(True)
    lg = logits.to(DEV).requires_grad_(True)
    assert MSDeformAttnFusedFunction.supported(v, ref.to(DEV), len(shapes), P)
    out = MSDeformAttnFusedFunction.apply(v, sh.to(DEV), lsi.to(DEV), ref.to(DEV), o, lg)
    out.backward(gout.to(DEV, dtype))
    torch.cuda.synchronize()
    r_out, r_gv, r_goff, r_glog = reference(value, sh, lsi, ref, off, logits, gout, norm, P)
    ft, bt = (1e-5, 1e-4) if dtype == torch.float32 else (2e-2, 2e-2)
    errs = dict(out=rel(out, r_out), gv=rel(v.grad, r_gv), goff=rel(o.grad, r_goff), glog=rel(lg.grad, r_glog))
    assert errs["out"] <= ft and errs["gv"] <= bt and errs["goff"] <= bt and errs["glog"] <= bt, errs


def test_fused_equals_unfused_at_benchmark_size():
    """Full CVPPP-512 pyramid: fused route vs torch prologue + MSDeformAttnFunction (both CUDA)."""
    shapes, M, D, P = synth.CVPPP_512, 8, 32, 4
    value, sh, lsi, ref, off, logits, gout, norm = make_case(4, shapes, M, D, P, 2, seed=7)
    v1 = value.to(DEV).requires_grad_(True)
    o1 = off.to(DEV).requires_grad_(True)
    l1 = logits.to(DEV).requires_grad_(True)
    out1 = MSDeformAttnFusedFunction.apply(v1, sh.to(DEV), lsi.to(DEV), ref.to(DEV), o1, l1)
    out1.backward(gout.to(DEV))
    v2 = value.to(DEV).requires_grad_(True)
    o2 = off.to(DEV).requires_grad_(True)
    l2 = logits.to(DEV).requires_grad_(True)
    loc = ref.to(DEV)[:, :, None, :, None, :] + o2 / norm.to(DEV)[None, None, None, :, None, :]
    attn = torch.softmax(l2, -1).view(4, -1, M, len(shapes), P)
    out2 = MSDeformAttnFunction.apply(v2, sh.to(DEV), lsi.to(DEV), loc.contiguous(), attn.contiguous(), 128)
    out2.backward(gout.to(DEV))
    assert rel(out1, out2) <= 1e-5
    assert rel(v1.grad, v2.grad) <= 1e-4 and rel(o1.grad, o2.grad) <= 1e-4 and rel(l1.grad, l2.grad) <= 1e-4


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_fused_vs_unfused(ref_dim):
    torch.manual_seed(0)
    m = MSDeformAttn(d_model=128, n_levels=3, n_heads=8, n_points=4).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.sampling_offsets.bias.add_(0.31)
        m.attention_weights.weight.normal_(0, 0.1)
    shapes = ((12, 10), (6, 5), (3, 3))
    sh, lsi, S = synth.level_tensors(shapes, device=DEV)
    q = torch.randn(2, 50, 128, device=DEV)
    src = torch.randn(2, S, 128, device=DEV)
    ref = torch.rand(2, 50, 3, ref_dim, device=DEV) * 0.8 + 0.1
    mask = torch.zeros(2, S, dtype=torch.bool, device=DEV)
    mask[1, -5:] = True
    outs, grads = [], []
    for fused in (True, False):
        m.fused_prologue = fused
        m.zero_grad(set_to_none=True)
        out = m(q, ref, src, sh, lsi, mask)
        out.square().sum().backward()
        outs.append(out.detach())
        grads.append({k: p.grad.clone() for k, p in m.named_parameters()})
    assert rel(outs[0], outs[1]) <= 1e-5
    for k in grads[0]:
        assert rel(grads[0][k], grads[1][k]) <= 2e-4, k


def test_fused_unsupported_shapes_are_reported():
    v = torch.randn(1, 6, 2, 30, device=DEV)           # D=30: generic kernels only
    ref = torch.rand(1, 4, 1, 2, device=DEV)
    assert not MSDeformAttnFusedFunction.supported(v, ref, 1, 2)
    assert not MSDeformAttnFusedFunction.supported(v.double(), ref, 1, 2)
    sh = torch.tensor([[2, 3]], device=DEV)
    lsi = torch.zeros(1, dtype=torch.long, device=DEV)
    with pytest.raises(RuntimeError):
        MSDeformAttnFusedFunction.apply(v, sh, lsi, ref, torch.zeros(1, 4, 2, 1, 2, 2, device=DEV),
                                        torch.zeros(1, 4, 2, 2, device=DEV))


def test_module_under_bf16_autocast_tracks_fp32():
    """The reference pins this block to fp32 because its kernels have no half path
    (PD/msdeformattn.py:314-320, cu:69); ours accepts bf16 values: under autocast the module's
    output stays within bf16 accuracy of the fp32 run, forward and parameter gradients."""
    torch.manual_seed(1)
    m = MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.sampling_offsets.bias.add_(0.27)
        m.attention_weights.weight.normal_(0, 0.1)
    shapes = ((16, 16), (8, 8), (4, 4))
    sh, lsi, S = synth.level_tensors(shapes, device=DEV)
    src = torch.randn(2, S, 256, device=DEV)
    ref = synth.encoder_reference_points(shapes).to(DEV)[None, :, None, :].expand(2, S, 3, 2).contiguous()

    def run(autocast):
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = m(src, ref, src, sh, lsi)
        out.float().square().mean().backward()
        return out.detach().float(), m.value_proj.weight.grad.clone(), m.sampling_offsets.weight.grad.clone()

    o32, gv32, go32 = run(False)
    o16, gv16, go16 = run(True)
    assert rel(o16, o32) <= 3e-2
    assert float((gv16 - gv32).norm() / gv32.norm()) <= 5e-2
    # d/d(location) is a difference of neighbouring bf16-rounded values: the noisiest gradient
    assert float((go16 - go32).norm() / go32.norm()) <= 0.25
