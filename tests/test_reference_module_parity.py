"""CPU, authoring container only: our MSDeformAttn module against the REFERENCE's own module class
(imported from /root/reference with a stub for its un-built extension).  Skipped where the
reference tree does not exist (the GPU box)."""
import importlib
import os
import sys
import types

import pytest
import torch

OPS_PARENT = "/root/reference/connectomics/model/maskformer_block/pixel_decoder"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(OPS_PARENT, "ops")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_module_cls():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, OPS_PARENT)
    try:
        mod = importlib.import_module("ops.modules.ms_deform_attn")
    finally:
        sys.path.remove(OPS_PARENT)
    return mod.MSDeformAttn


@pytest.mark.parametrize("cfg", [(256, 3, 8, 4), (128, 3, 8, 4), (64, 4, 4, 2)])
def test_same_parameters_after_init(ref_module_cls, cfg):
    from pctrans_b200.modules import MSDeformAttn
    torch.manual_seed(7)
    ref = ref_module_cls(*cfg)
    torch.manual_seed(7)
    ours = MSDeformAttn(*cfg)
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r) == list(sd_o)
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), k
    assert ours.im2col_step == ref.im2col_step


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_same_forward_as_reference_module_on_cpu(ref_module_cls, monkeypatch, ref_dim):
    """The reference module on CPU takes its own PyTorch path (its try/except, mod.py:119-121); ours
    is pointed at the oracle for this test only.  Same weights, same inputs -> same output."""
    from oracle import msda_oracle as oracle
    from pctrans_b200 import synth
    from pctrans_b200.modules import MSDeformAttn, ms_deform_attn as mod_file

    class _OracleFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, step):
            return oracle.core_grid_sample(value, shapes, loc, attn)

    monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFunction)
    torch.manual_seed(3)
    ref = ref_module_cls(64, 3, 4, 4)
    ours = MSDeformAttn(64, 3, 4, 4)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    ours.load_state_dict(ref.state_dict())
    shapes = ((8, 6), (4, 3), (2, 2))
    sh, lsi, S = synth.level_tensors(shapes)
    q = torch.randn(2, 20, 64)
    src = torch.randn(2, S, 64)
    rp = torch.rand(2, 20, 3, ref_dim) * 0.8 + 0.1
    mask = torch.zeros(2, S, dtype=torch.bool)
    mask[0, :3] = True
    with torch.no_grad():
        want = ref(q, rp, src, sh, lsi, mask)
        got = ours(q, rp, src, sh, lsi, mask)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.fixture(scope="module")
def ref_encoder_module(ref_module_cls):
    """The reference's pixel_decoder/msdeformattn.py with detectron2 / fvcore replaced by inert
    stubs (only the pure-torch encoder classes are used) and the sibling packages registered bare
    so their __init__ (which needs removed torch internals) does not run."""

    class _Meta(type):
        def __getattr__(cls, name):
            return _Any

        def __call__(cls, *a, **k):
            if len(a) == 1 and not k and callable(a[0]):
                return a[0]
            return type.__call__(cls)

    class _Any(metaclass=_Meta):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return a[0] if (len(a) == 1 and not k and callable(a[0])) else self

        def __getattr__(self, name):
            return _Any

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Any

    for name in ("detectron2", "detectron2.config", "detectron2.layers", "detectron2.modeling",
                 "fvcore", "fvcore.nn", "fvcore.nn.weight_init"):
        if name not in sys.modules:
            stub = _Stub(name)
            stub.__path__ = []
            sys.modules[name] = stub
    root = "/root/reference/connectomics/model/maskformer_block"
    pkg = types.ModuleType("_ref_mbk")
    pkg.__path__ = [root]
    sys.modules["_ref_mbk"] = pkg
    for sub in ("transformer_decoder", "pixel_decoder"):
        p = types.ModuleType("_ref_mbk." + sub)
        p.__path__ = [os.path.join(root, sub)]
        sys.modules["_ref_mbk." + sub] = p
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("_ref_mbk.pixel_decoder.msdeformattn")


def test_encoder_matches_reference_encoder(ref_encoder_module, monkeypatch):
    """MSDeformAttnTransformerEncoderOnly + PositionEmbeddingSine: identical parameters after a
    seeded init and identical memory / shapes / level_start_index on CPU."""
    from oracle import msda_oracle as oracle
    from pctrans_b200 import encoder as enc
    from pctrans_b200.modules import ms_deform_attn as mod_file

    class _OracleFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, step):
            return oracle.core_grid_sample(value, shapes, loc, attn)

    monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFunction)
    kw = dict(d_model=64, nhead=4, num_encoder_layers=2, dim_feedforward=128, dropout=0.0,
              num_feature_levels=3)
    torch.manual_seed(11)
    ref = ref_encoder_module.MSDeformAttnTransformerEncoderOnly(**kw).eval()
    torch.manual_seed(11)
    ours = enc.MSDeformAttnTransformerEncoderOnly(**kw).eval()
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r) == list(sd_o)
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), k

    srcs = [torch.randn(2, 64, h, w) for h, w in ((3, 4), (6, 8), (12, 16))]
    pe_r = ref_encoder_module.PositionEmbeddingSine(32, normalize=True)
    pe_o = enc.PositionEmbeddingSine(32, normalize=True)
    pos_r, pos_o = [pe_r(s) for s in srcs], [pe_o(s) for s in srcs]
    for a, b in zip(pos_r, pos_o):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    with torch.no_grad():
        mem_r, sh_r, lsi_r = ref(srcs, pos_r)
        mem_o, sh_o, lsi_o = ours(srcs, pos_o)
    assert torch.equal(sh_r, sh_o) and torch.equal(lsi_r, lsi_o)
    assert torch.allclose(mem_o, mem_r, rtol=1e-5, atol=1e-5)


def test_pixel_decoder_matches_reference_decoder(ref_encoder_module, monkeypatch):
    """MSDeformAttnPixelDecoder (PD:164-360): the reference class is instantiated with minimal
    real stand-ins for the four detectron2 / fvcore helpers it needs (Conv2d-with-norm, get_norm,
    ShapeSpec, c2_xavier_fill); its state_dict must load into ours with strict=True (checkpoint
    compatibility) and forward_features must agree on CPU."""
    import collections
    import torch.nn.functional as F
    from torch import nn
    from oracle import msda_oracle as oracle
    from pctrans_b200 import encoder as enc
    from pctrans_b200.modules import ms_deform_attn as mod_file

    class Conv2d(nn.Conv2d):                       # detectron2.layers.Conv2d semantics
        def __init__(self, *a, norm=None, activation=None, **k):
            super().__init__(*a, **k)
            self.norm, self.activation = norm, activation

        def forward(self, x):
            x = super().forward(x)
            if self.norm is not None:
                x = self.norm(x)
            return self.activation(x) if self.activation is not None else x

    ShapeSpec = collections.namedtuple("ShapeSpec", ["channels", "stride"])

    def c2_xavier_fill(m):
        nn.init.kaiming_uniform_(m.weight, a=1)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)

    ref_mod = ref_encoder_module
    monkeypatch.setattr(ref_mod, "Conv2d", Conv2d)
    monkeypatch.setattr(ref_mod, "get_norm", lambda norm, c: nn.GroupNorm(32, c) if norm == "GN" else None)
    monkeypatch.setattr(ref_mod.weight_init, "c2_xavier_fill", c2_xavier_fill, raising=False)

    class _OracleFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, step):
            return oracle.core_grid_sample(value, shapes, loc, attn)

    monkeypatch.setattr(mod_file, "MSDeformAttnFunction", _OracleFunction)
    chans = {"res2": (32, 4), "res3": (48, 8), "res4": (64, 16), "res5": (96, 32)}
    torch.manual_seed(5)
    ref = ref_mod.MSDeformAttnPixelDecoder(
        {k: ShapeSpec(*v) for k, v in chans.items()}, transformer_dropout=0.0, transformer_nheads=4,
        transformer_dim_feedforward=128, transformer_enc_layers=2, conv_dim=64, mask_dim=64, norm="GN",
        transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    ours = enc.MSDeformAttnPixelDecoder(
        chans, transformer_dropout=0.0, transformer_nheads=4, transformer_dim_feedforward=128,
        transformer_enc_layers=2, conv_dim=64, mask_dim=64, norm="GN",
        transformer_in_features=("res3", "res4", "res5"), common_stride=4).eval()
    missing = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    g = torch.Generator().manual_seed(6)
    feats = {k: torch.randn(2, c, 64 // s, 96 // s, generator=g) for k, (c, s) in chans.items()}
    with torch.no_grad():
        r_mask, r_low, r_multi = ref.forward_features(feats)
        o_mask, o_low, o_multi = ours.forward_features(feats)
    assert torch.allclose(o_mask, r_mask, rtol=1e-4, atol=1e-5)
    assert torch.allclose(o_low, r_low, rtol=1e-4, atol=1e-5)
    assert len(o_multi) == len(r_multi) == 3
    for a, b in zip(o_multi, r_multi):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
