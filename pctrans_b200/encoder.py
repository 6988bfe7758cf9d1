"""Callers of the op: the deformable-attention encoder of PCTrans's pixel decoder, restated without
detectron2 / fvcore so BASELINE configs 2 and 4 can run.

Mirrors connectomics/model/maskformer_block/pixel_decoder/msdeformattn.py (PD):
  MSDeformAttnTransformerEncoderLayer  PD:92-131     MSDeformAttnTransformerEncoder  PD:134-161
  MSDeformAttnTransformerEncoderOnly   PD:23-89      MSDeformAttnPixelDecoder        PD:164-360
  PositionEmbeddingSine                transformer_decoder/position_encoding.py:12-52
Class names, constructor arguments and parameter names are kept so checkpoints
(``...pixel_decoder.transformer.encoder.layers.{i}.self_attn.*`` etc.) load unchanged.  Everything
except ``MSDeformAttn`` is stock PyTorch (cuBLAS GEMMs, LayerNorm, conv): those are library calls
outside the hot path.
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .functions import linear_tc
from .modules import MSDeformAttn


class PositionEmbeddingSine(nn.Module):
    """Sine/cosine position embedding over (y, x) of an image (position_encoding.py:12-52)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, x, mask=None):
        n, _, h, w = x.shape
        if mask is None:
            mask = torch.zeros((n, h, w), device=x.device, dtype=torch.bool)
        valid = ~mask
        ys = valid.cumsum(1, dtype=torch.float32)
        xs = valid.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            ys = ys / (ys[:, -1:, :] + eps) * self.scale
            xs = xs / (xs[:, :, -1:] + eps) * self.scale
        k = torch.arange(self.num_pos_feats, dtype=torch.float32, device=x.device)
        freq = self.temperature ** (2 * torch.div(k, 2, rounding_mode="floor") / self.num_pos_feats)

        def embed(coord):
            ang = coord[..., None] / freq
            return torch.stack((ang[..., 0::2].sin(), ang[..., 1::2].cos()), dim=4).flatten(3)

        return torch.cat((embed(ys), embed(xs)), dim=3).permute(0, 3, 1, 2)


def _activation(name: str):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu/glu, not {name}.")


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4,
                 n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        linear_tc.install_invalidation_hooks(self)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _fusable(self, *dropouts):
        """The Linear epilogues (ReLU, residual add, LayerNorm) can be fused into the tensor-core
        GEMMs when no dropout sits between them (p = 0 or eval mode)."""
        return self.self_attn.tensor_core_linears and all(d.p == 0.0 or not self.training for d in dropouts)

    def forward_ffn(self, src):
        if self._fusable(self.dropout2, self.dropout3) and self.activation is F.relu:
            return linear_tc.ffn_block(src, self.linear1, self.linear2, self.norm2)
        hidden = self.dropout2(self.activation(self.linear1(src)))
        return self.norm2(src + self.dropout3(self.linear2(hidden)))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        # query = src + pos, value source = src (PD:124)
        query = self.with_pos_embed(src, pos)
        if self._fusable(self.dropout1):
            src = self.self_attn(query, reference_points, src, spatial_shapes, level_start_index,
                                 padding_mask, residual=src, norm=self.norm1)
        else:
            attn = self.self_attn(query, reference_points, src, spatial_shapes, level_start_index,
                                  padding_mask)
            src = self.norm1(src + self.dropout1(attn))
        return self.forward_ffn(src)


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """(N, S, L, 2): normalised pixel centres of every pyramid position, replicated to all
        levels and scaled by the valid ratios (PD:141-153).  ``spatial_shapes`` may be the (L, 2)
        tensor (as in the reference; reading it costs a device sync) or a sequence of (H, W)."""
        per_level = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            ys = torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            gy = gy.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            gx = gx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            per_level.append(torch.stack((gx, gy), -1))
        ref = torch.cat(per_level, 1)
        return ref[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                *, shapes_hint=None, unit_ratios=False):
        """``shapes_hint``: the pyramid as Python (H, W) pairs when the caller has them, so the
        reference points are built without reading the shape tensor back from the device.
        ``unit_ratios``: the caller guarantees ``valid_ratios == 1`` (no padding), which makes the
        reference points a constant of the pyramid: they are built once and reused."""
        out = src
        if shapes_hint is not None and unit_ratios:
            cache = self.__dict__.setdefault("_ref_cache", {})
            key = (tuple(shapes_hint), src.shape[0], str(src.device))
            if key not in cache:
                # built outside inference mode: a cached inference tensor could not be saved for
                # backward by a later training step
                with torch.inference_mode(False):
                    cache[key] = self.get_reference_points(
                        shapes_hint, torch.ones_like(valid_ratios), device=src.device)
            ref = cache[key]
        else:
            ref = self.get_reference_points(spatial_shapes if shapes_hint is None else shapes_hint,
                                            valid_ratios, device=src.device)
        for layer in self.layers:
            out = layer(out, pos, ref, spatial_shapes, level_start_index, padding_mask)
        return out


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024,
                 dropout=0.1, activation="relu", num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                    num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        linear_tc.invalidate()   # the initialisers below may write through .data
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def _level_tensors(self, shapes, device):
        """(L, 2) shapes and (L,) start indices on the device, built once per pyramid: the host ->
        device copy of a Python list each forward would be a sync point (and is not capturable in
        a CUDA graph)."""
        cache = self.__dict__.setdefault("_level_cache", {})
        key = (shapes, str(device))
        if key not in cache:
            with torch.inference_mode(False):   # see _ref_cache: never cache inference tensors
                spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=device)
                level_start_index = torch.cat(
                    (spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
            if spatial_shapes.is_cuda:
                # the backward kernels may use the exact cell count of the pyramid; the shapes are
                # known here on the host, so the binding never has to read the tensor back
                from . import _cabi
                _cabi.register_host_shapes(spatial_shapes, shapes)
            cache[key] = (spatial_shapes, level_start_index)
        return cache[key]

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_h = torch.sum(~mask[:, :, 0], 1).float() / H
        valid_w = torch.sum(~mask[:, 0, :], 1).float() / W
        return torch.stack([valid_w, valid_h], -1)

    def forward(self, srcs: Sequence[torch.Tensor], pos_embeds: Sequence[torch.Tensor]):
        """srcs / pos_embeds: per level (N, C, H_l, W_l), low -> high resolution.  Returns
        (memory (N, S, C), spatial_shapes (L, 2), level_start_index (L,))."""
        # No padding anywhere (the reference builds all-False masks here, PD:62, and derives
        # valid_ratios == 1 from them): the masks are not materialised, the ratios are ones.
        flat_src, flat_pos, shapes = [], [], []
        for lvl, (src, pos) in enumerate(zip(srcs, pos_embeds)):
            n, _, h, w = src.shape
            shapes.append((h, w))
            flat_src.append(src.flatten(2).transpose(1, 2))
            flat_pos.append(pos.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(flat_src, 1)
        pos_flatten = torch.cat(flat_pos, 1)
        spatial_shapes, level_start_index = self._level_tensors(tuple(shapes), src_flatten.device)
        valid_ratios = src_flatten.new_ones((src_flatten.shape[0], len(shapes), 2))
        # ... and the layers are told there is no padding instead of running masked_fill over every
        # value with an all-False mask
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios,
                              pos_flatten, None, shapes_hint=shapes, unit_ratios=True)
        return memory, spatial_shapes, level_start_index


class _ConvNorm(nn.Conv2d):
    """Conv2d carrying an optional ``norm`` sub-module and activation, so state-dict keys match
    detectron2.layers.Conv2d (``adapter_1.weight``, ``adapter_1.norm.weight`` ...)."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _c2_xavier_fill(conv: nn.Conv2d):
    nn.init.kaiming_uniform_(conv.weight, a=1)   # fvcore.nn.weight_init.c2_xavier_fill
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)


class MSDeformAttnPixelDecoder(nn.Module):
    """PD:164-360 without the detectron2 registry / ``configurable`` plumbing.  ``input_shape`` maps
    feature name -> (channels, stride), e.g. {"res2": (256, 4), ..., "res5": (2048, 32)}."""

    def __init__(self, input_shape: Dict[str, tuple], *, transformer_dropout: float = 0.0,
                 transformer_nheads: int = 8, transformer_dim_feedforward: int = 1024,
                 transformer_enc_layers: int = 6, conv_dim: int = 256, mask_dim: int = 256,
                 norm: str = "GN",
                 transformer_in_features: Sequence[str] = ("res3", "res4", "res5"),
                 common_stride: int = 4):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1][1])
        self.in_features = [k for k, _ in by_stride]
        self.feature_channels = [v[0] for _, v in by_stride]
        tf = [(k, v) for k, v in by_stride if k in transformer_in_features]
        self.transformer_in_features = [k for k, _ in tf]
        self.transformer_feature_strides = [v[1] for _, v in tf]
        self.transformer_num_feature_levels = len(tf)
        # low resolution first (res5 -> res3), PD:213-219
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(v[0], conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim))
            for _, v in tf[::-1]])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)
        self.mask_dim = mask_dim
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        self.num_fpn_levels = int(math.log2(min(self.transformer_feature_strides)) - math.log2(common_stride))
        lateral, output = [], []
        use_bias = norm == ""
        make_norm = (lambda: nn.GroupNorm(32, conv_dim)) if norm == "GN" else (lambda: None)
        for idx, in_ch in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lat = _ConvNorm(in_ch, conv_dim, kernel_size=1, bias=use_bias, norm=make_norm())
            out = _ConvNorm(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                            norm=make_norm(), activation=F.relu)
            _c2_xavier_fill(lat)
            _c2_xavier_fill(out)
            self.add_module(f"adapter_{idx + 1}", lat)
            self.add_module(f"layer_{idx + 1}", out)
            lateral.append(lat)
            output.append(out)
        self.lateral_convs = lateral[::-1]
        self.output_convs = output[::-1]

    @classmethod
    def from_config(cls, cfg, input_shape):
        """Constructor arguments from a PCTrans / Mask2Former config node (PD:288-304), without
        detectron2's ``configurable`` plumbing: ``MSDeformAttnPixelDecoder(**from_config(cfg, shapes))``.
        ``cfg`` is any object with the ``MODEL.SEM_SEG_HEAD.*`` / ``MODEL.MASK_FORMER.*`` attributes
        (a yacs / detectron2 CfgNode, a namespace); ``input_shape`` maps feature name to either a
        ``(channels, stride)`` pair or an object with ``.channels`` / ``.stride`` (ShapeSpec)."""
        head, mf = cfg.MODEL.SEM_SEG_HEAD, cfg.MODEL.MASK_FORMER

        def pair(v):
            return (int(v.channels), int(v.stride)) if hasattr(v, "channels") else (int(v[0]), int(v[1]))

        return {
            "input_shape": {k: pair(v) for k, v in input_shape.items() if k in head.IN_FEATURES},
            "conv_dim": head.CONVS_DIM,
            "mask_dim": head.MASK_DIM,
            "norm": head.NORM,
            "transformer_dropout": mf.DROPOUT,
            "transformer_nheads": mf.NHEADS,
            "transformer_dim_feedforward": 1024,   # fixed for the deformable encoder (PD:298-299)
            "transformer_enc_layers": head.TRANSFORMER_ENC_LAYERS,
            "transformer_in_features": head.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES,
            "common_stride": head.COMMON_STRIDE,
        }

    def forward_features(self, features: Dict[str, torch.Tensor]):
        """Returns (mask_features, lowest-resolution encoder map, multi_scale_features[:3]) as
        PD:314-360; always fp32 (PD:314,320)."""
        with torch.autocast(device_type="cuda", enabled=False):
            srcs, pos = [], []
            for idx, name in enumerate(self.transformer_in_features[::-1]):
                x = features[name].float()
                srcs.append(self.input_proj[idx](x))
                pos.append(self.pe_layer(x))
            memory, shapes, lsi = self.transformer(srcs, pos)
            n = memory.shape[0]
            sizes = [int(h) * int(w) for h, w in shapes.tolist()]
            out: List[torch.Tensor] = [
                z.transpose(1, 2).reshape(n, -1, int(h), int(w))
                for z, (h, w) in zip(memory.split(sizes, dim=1), shapes.tolist())]
            for idx, name in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                lat = self.lateral_convs[idx](features[name].float())
                # PCTrans resamples the lateral map onto the encoder map's grid (PD:349), unlike
                # upstream Mask2Former which upsamples the encoder map
                y = out[-1] + F.interpolate(lat, size=out[-1].shape[-2:], mode="bilinear",
                                            align_corners=False)
                out.append(self.output_convs[idx](y))
            return out[-1], out[0], out[:self.maskformer_num_feature_levels]
