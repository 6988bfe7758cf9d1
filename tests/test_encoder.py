"""The callers of the op (pixel-decoder encoder, pctrans_b200/encoder.py) -- BASELINE configs 2
and 4.  The same module weights are run twice: once through the CUDA op and once with the op
replaced by the oracle's grid_sample port (monkeypatched in the TEST only; the product has no
such switch)."""
import pytest
import torch

from oracle import msda_oracle as oracle
from pctrans_b200 import encoder as enc
from pctrans_b200.modules import ms_deform_attn as mod_file


class _OracleFunction:
    """Stands in for MSDeformAttnFunction inside tests: same call signature, oracle arithmetic."""

    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return oracle.core_grid_sample(value, shapes, loc, attn)


class _OracleFp32Function(torch.autograd.Function):
    """MSDeformAttnFunction with the C oracle as arithmetic, in the dtype of ``value``: the literal
    fp32 instantiation rounds the pixel coordinate exactly like the kernels (cuh:290-291), so both
    legs sample the SAME bilinear cells even for points within rounding distance of a grid line --
    which a grid_sample formulation (different coordinate arithmetic) does not guarantee; with
    fp64 tensors it is the ground truth both fp32 evaluations are measured against."""

    @staticmethod
    def forward(ctx, value, shapes, lsi, loc, attn, im2col_step):
        import numpy as np
        ctx.save_for_backward(value, shapes, lsi, loc, attn)
        ctx.np_dtype = np.float64 if value.dtype == torch.float64 else np.float32
        out = oracle.forward_c(value, shapes, lsi, loc, attn, dtype=ctx.np_dtype)
        return torch.from_numpy(out).to(value.device)

    @staticmethod
    def backward(ctx, grad_out):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = oracle.backward_c(value, shapes, lsi, loc, attn, grad_out.contiguous(), dtype=ctx.np_dtype)
        dev = value.device
        return (torch.from_numpy(gv).to(dev), None, None, torch.from_numpy(gl).to(dev),
                torch.from_numpy(ga).to(dev), None)


def _set_fused(model, on):
    """CUDA leg: fused-prologue kernels; oracle leg: the module composes softmax / locations in
    PyTorch and calls the (monkeypatched) Function."""
    from pctrans_b200.modules import MSDeformAttn
    for m in model.modules():
        if isinstance(m, MSDeformAttn):
            m.fused_prologue = on


def _features(n, sizes, chans=(256, 512, 1024, 2048), seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return {f"res{i + 2}": torch.randn(n, c, h, w, generator=g).to(device)
            for i, (c, (h, w)) in enumerate(zip(chans, sizes))}


def _decoder(conv_dim, seed=0, layers=6):
    torch.manual_seed(seed)
    return enc.MSDeformAttnPixelDecoder(
        {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)},
        conv_dim=conv_dim, mask_dim=conv_dim, transformer_enc_layers=layers, transformer_dropout=0.0)


def test_state_dict_keys_match_reference_layout():
    dec = _decoder(128, layers=2)
    keys = set(dec.state_dict())
    for k in ("transformer.level_embed",
              "transformer.encoder.layers.1.self_attn.sampling_offsets.bias",
              "transformer.encoder.layers.0.self_attn.value_proj.weight",
              "transformer.encoder.layers.0.norm1.weight", "transformer.encoder.layers.0.linear1.weight",
              "transformer.encoder.layers.0.linear2.bias", "transformer.encoder.layers.0.norm2.bias",
              "input_proj.0.0.weight", "input_proj.2.1.bias", "adapter_1.weight",
              "adapter_1.norm.weight", "layer_1.weight", "layer_1.norm.bias"):
        assert k in keys, k
    # d256 encoder parameter count quoted in SURVEY.md 2a (4 393 920 incl. level_embed)
    assert sum(p.numel() for p in _decoder(256).transformer.parameters()) == 4_393_920


def test_from_config_reads_the_pctrans_config_node():
    """PD:288-304 with the values of configs/CVPPP/CVPPP-PCTrans.yaml (CONVS_DIM 128 etc.)."""
    from types import SimpleNamespace as NS
    cfg = NS(MODEL=NS(
        SEM_SEG_HEAD=NS(IN_FEATURES=["res2", "res3", "res4", "res5"], CONVS_DIM=128, MASK_DIM=128, NORM="GN",
                        TRANSFORMER_ENC_LAYERS=6, COMMON_STRIDE=4,
                        DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"]),
        MASK_FORMER=NS(DROPOUT=0.0, NHEADS=8, DIM_FEEDFORWARD=2048)))
    shapes = {"res2": NS(channels=256, stride=4), "res3": (512, 8), "res4": (1024, 16),
              "res5": NS(channels=2048, stride=32), "stem": (64, 2)}
    kw = enc.MSDeformAttnPixelDecoder.from_config(cfg, shapes)
    assert kw["transformer_dim_feedforward"] == 1024 and "stem" not in kw["input_shape"]
    dec = enc.MSDeformAttnPixelDecoder(**kw)
    assert dec.transformer.d_model == 128 and len(dec.transformer.encoder.layers) == 6
    assert dec.transformer_in_features == ["res3", "res4", "res5"] and dec.num_fpn_levels == 1


def test_decoder_runs_on_cpu_through_oracle(monkeypatch):
    """Host logic only (no GPU): shapes, level order low->high resolution, FPN output grid."""
    monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFunction)
    dec = _decoder(64, layers=1).eval()
    feats = _features(1, ((32, 32), (16, 16), (8, 8), (4, 4)))
    with torch.no_grad():
        mask_feat, lowest, multi = dec.forward_features(feats)
        srcs = [dec.input_proj[i](feats[f]) for i, f in enumerate(("res5", "res4", "res3"))]
        pos = [dec.pe_layer(feats[f]) for f in ("res5", "res4", "res3")]
        memory, shapes, lsi = dec.transformer(srcs, pos)
    assert shapes.tolist() == [[4, 4], [8, 8], [16, 16]] and lsi.tolist() == [0, 16, 80]
    assert memory.shape == (1, 336, 64)
    assert lowest.shape == (1, 64, 4, 4) and mask_feat.shape == (1, 64, 16, 16)
    assert [m.shape[-1] for m in multi] == [4, 8, 16]
    ref = enc.MSDeformAttnTransformerEncoder.get_reference_points(
        [(2, 3)], torch.ones(1, 1, 2), "cpu")
    assert torch.allclose(ref[0, 0, 0], torch.tensor([0.5 / 3, 0.25]))


def test_module_without_cuda_fails_loudly():
    dec = _decoder(64, layers=1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        dec.forward_features(_features(1, ((16, 16), (8, 8), (4, 4), (2, 2))))


@pytest.mark.gpu
@pytest.mark.parametrize("conv_dim", [256, 128])
def test_config2_decoder_forward_matches_oracle(monkeypatch, conv_dim):
    """BASELINE config 2: pixel-decoder forward (inference), batch 8, 512x512 pyramid; the d128
    variant is PCTrans's real CONVS_DIM (head dim 16)."""
    dev = "cuda:0"
    dec = _decoder(conv_dim).to(dev).eval()
    feats = _features(8, ((128, 128), (64, 64), (32, 32), (16, 16)), device=dev)
    with torch.no_grad():
        got = dec.forward_features(feats)
        srcs = [dec.input_proj[i](feats[f]) for i, f in enumerate(("res5", "res4", "res3"))]
        pos = [dec.pe_layer(feats[f]) for f in ("res5", "res4", "res3")]
        mem = dec.transformer(srcs, pos)[0]
        monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFunction)
        _set_fused(dec, False)
        want = dec.forward_features(feats)
        mem_ref = dec.transformer(srcs, pos)[0]
    assert mem.shape == (8, 5376, conv_dim)
    err = float((mem - mem_ref).abs().max() / mem_ref.abs().max())
    assert err <= 1e-5, err            # accumulates over 6 layers
    # the FPN conv + GroupNorm after the encoder amplifies the 1e-5 differences a little
    for a, b in zip((got[0], got[1], *got[2]), (want[0], want[1], *want[2])):
        assert float((a - b).abs().max() / b.abs().max()) <= 2e-4


@pytest.mark.gpu
def test_config4_bbbc_encoder_forward_backward(monkeypatch):
    """BASELINE config 4 shape (520x696 pyramid, 6 encoder layers, fwd+bwd), small batch for the
    oracle leg; parameter gradients of the two legs must agree."""
    dev = "cuda:0"
    torch.manual_seed(0)
    model = enc.MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=6,
                                                   dim_feedforward=1024, dropout=0.0,
                                                   num_feature_levels=3).to(dev)
    # Freshly initialised offsets are whole pixels (mod.py:66-74), which puts every sample ON the
    # pixel grid where d/d(location) is one-sided and two fp32 formulations may pick different
    # cells; give the layers "trained-like" generic offsets and non-uniform attention instead.
    with torch.no_grad():
        for layer in model.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.self_attn.sampling_offsets.bias.add_(
                torch.rand_like(layer.self_attn.sampling_offsets.bias) * 0.6 + 0.2)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    g = torch.Generator().manual_seed(3)
    sizes = ((17, 22), (33, 44), (65, 87))
    srcs = [torch.randn(2, 256, h, w, generator=g).to(dev) for h, w in sizes]
    pe = enc.PositionEmbeddingSine(128, normalize=True)
    pos = [pe(s) for s in srcs]

    def run():
        model.zero_grad(set_to_none=True)
        mem, shapes, _ = model(srcs, pos)
        mem.square().mean().backward()
        return mem.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    mem_fused, grads_fused = run()          # product path: fused-prologue kernels
    _set_fused(model, False)
    mem, grads = run()                      # the unfused Function (locations / weights from PyTorch)
    monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFp32Function)
    mem_ref, grads_ref = run()              # same module code, op = literal-fp32 C oracle (same cells)
    assert mem.shape == (2, 7481, 256)
    assert float((mem - mem_ref).abs().max() / mem_ref.abs().max()) <= 1e-5
    assert float((mem_fused - mem_ref).abs().max() / mem_ref.abs().max()) <= 1e-5

    # Ground truth: the same encoder in fp64 (fp64 GEMMs, fp64 C oracle as the op).  Parameter
    # gradients are sums over ~15 000 positions with heavy cancellation, so two fp32 evaluations
    # differ from each other by far more than 1e-4 on some tensors (sampling_offsets.weight of the
    # late layers: ~1e-2) -- the honest bar is that the CUDA path is as close to the truth as a
    # literal fp32 evaluation of the reference's own formulas, tensor by tensor.
    model.double()
    srcs, pos = [s.double() for s in srcs], [p.double() for p in pos]
    _, grads64 = run()

    def errors(got):
        out = {}
        for k, ref in grads64.items():
            scale = float(ref.abs().max())
            if scale == 0.0:
                assert float(got[k].abs().max()) == 0.0, k
                continue
            out[k] = float((got[k].double() - ref).norm() / ref.norm())
        return out

    e_ref32, e_cuda, e_fused = errors(grads_ref), errors(grads), errors(grads_fused)
    for k in e_ref32:
        bar = max(1e-3, 4.0 * e_ref32[k])
        assert e_cuda[k] <= bar, ("unfused", k, e_cuda[k], e_ref32[k])
        assert e_fused[k] <= bar, ("fused", k, e_fused[k], e_ref32[k])
    worst = max(e_cuda, key=e_cuda.get)
    print(f"config 4 parameter gradients vs fp64: worst CUDA {worst} {e_cuda[worst]:.2e} "
          f"(literal fp32 oracle on the same tensor {e_ref32[worst]:.2e}); "
          f"median CUDA {sorted(e_cuda.values())[len(e_cuda) // 2]:.2e}")


def test_linear_helpers_on_cpu_are_plain_torch():
    """functions/linear_tc.py only takes over CUDA fp32 GEMMs; on CPU tensors its helpers are the
    torch expressions they replace (the encoder's CPU path, used with the oracle above)."""
    from pctrans_b200.functions import linear_tc
    torch.manual_seed(0)
    lin, norm = torch.nn.Linear(32, 48), torch.nn.LayerNorm(48)
    x, res = torch.randn(5, 7, 32), torch.randn(5, 7, 48)
    assert not linear_tc.eligible(x, lin.weight)
    assert torch.equal(linear_tc.linear(x, lin.weight, lin.bias), lin(x))
    assert torch.equal(linear_tc.linear(x, lin.weight, lin.bias, residual=res, relu=True),
                       torch.relu(lin(x) + res))
    assert torch.equal(linear_tc.linear_residual_layernorm(x, lin.weight, lin.bias, res, norm),
                       norm(res + lin(x)))
    mask = torch.zeros(5, 7, dtype=torch.bool)
    mask[2, 3:] = True
    assert torch.equal(linear_tc.linear(x, lin.weight, lin.bias, row_mask=mask),
                       lin(x).masked_fill(mask[..., None], 0.0))          # mod.py:100-101


@pytest.mark.gpu
def test_inference_mode_then_training():
    """The reference's module works under torch.inference_mode(); ours caches level / reference
    tensors across calls, and those caches must not hold inference tensors (they could not be
    saved for backward by a later training step)."""
    dev = "cuda:0"
    torch.manual_seed(0)
    model = enc.MSDeformAttnTransformerEncoderOnly(d_model=128, nhead=8, num_encoder_layers=2,
                                                   dim_feedforward=256, dropout=0.0,
                                                   num_feature_levels=3).to(dev)
    g = torch.Generator().manual_seed(1)
    srcs = [torch.randn(2, 128, h, w, generator=g).to(dev) for h, w in ((4, 5), (8, 10), (16, 20))]
    pe = enc.PositionEmbeddingSine(64, normalize=True)
    pos = [pe(s) for s in srcs]
    with torch.inference_mode():
        a = model(srcs, pos)[0]
    with torch.no_grad():
        b = model(srcs, pos)[0]
    assert torch.equal(a, b)
    out = model(srcs, pos)[0]                 # training step after the inference-mode call
    out.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.gpu
def test_packed_weight_cache_sees_data_writes_after_invalidate():
    """functions/linear_tc.py caches the packed 3xTF32 weights.  In-place updates under autograd
    are seen through the version counter; writes through ``.data`` are not -- the documented
    contract is ``linear_tc.invalidate()``, which the modules call from load_state_dict, train() /
    eval() and _reset_parameters."""
    from pctrans_b200.functions import linear_tc
    dev = "cuda:0"
    torch.manual_seed(0)
    layer = enc.MSDeformAttnTransformerEncoderLayer(d_model=64, d_ffn=128, dropout=0.0, n_levels=2,
                                                    n_heads=4, n_points=2).to(dev).eval()
    x = torch.randn(3, 40, 64, device=dev)
    with torch.no_grad():
        y0 = layer.forward_ffn(x)
        layer.linear1.weight.mul_(2.0)                    # version counter: seen
        y1 = layer.forward_ffn(x)
        assert not torch.allclose(y0, y1)
        ref = layer.norm2(x + layer.linear2(torch.relu(torch.nn.functional.linear(x, layer.linear1.weight, layer.linear1.bias))))
        assert torch.allclose(y1, ref, atol=1e-4)
        layer.linear1.weight.data.mul_(0.5)               # .data write: invisible to the cache ...
        linear_tc.invalidate()                            # ... until the documented call
        assert torch.allclose(layer.forward_ffn(x), y0, atol=1e-4)
        state = {k: v.clone() for k, v in layer.state_dict().items()}
        state["linear2.weight"] = state["linear2.weight"] * 3.0
        layer.load_state_dict(state)                      # hook invalidates
        ref = layer.norm2(x + layer.linear2(torch.relu(layer.linear1(x))))
        assert torch.allclose(layer.forward_ffn(x), ref, atol=1e-4)
        layer.linear2.weight.data.div_(3.0)
        layer.train().eval()                              # phase switch invalidates
        assert torch.allclose(layer.forward_ffn(x), y0, atol=1e-4)
