"""``MSDeformAttn`` -- the nn.Module around the B200 kernels.

Interface parity with the reference module (OPS/modules/ms_deform_attn.py:34-125) is deliberate:
constructor ``(d_model=256, n_levels=4, n_heads=8, n_points=4)``, the four Linear sub-modules
``sampling_offsets`` / ``attention_weights`` / ``value_proj`` / ``output_proj`` (so PCTrans
checkpoints load under ``...encoder.layers.{i}.self_attn.*``), ``_reset_parameters()`` (called
from outside, PD/msdeformattn.py:47-49), the ``im2col_step`` attribute and
``forward(query, reference_points, input_flatten, input_spatial_shapes,
input_level_start_index, input_padding_mask=None)``.

What differs on purpose:
  * no ``try/except`` around the op and no PyTorch fallback (the reference swallows every error
    and silently takes its slow path, mod.py:116-121): a failing CUDA call raises;
  * when the call qualifies, softmax and location arithmetic run inside the kernels
    (``fused_prologue``); otherwise they are composed here and the plain Function is used.  Both
    routes end in the same sm_100a kernels;
  * the four projections run on the tensor cores at fp32 accuracy when they qualify
    (``tensor_core_linears``, functions/linear_tc.py); ``forward`` takes two optional keyword
    arguments so that the encoder layer can have its residual add and LayerNorm fused into
    ``output_proj``.
"""
from __future__ import annotations

import math
import warnings

import weakref

import torch
from torch import nn

from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, linear_tc


def compass_offsets(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """Initial sampling-offset bias, flattened to (n_heads*n_levels*n_points*2,): head h looks along
    angle 2*pi*h/n_heads (scaled so the larger component is one pixel) and its k-th point sits k+1
    pixels out, identically on every level (mod.py:66-74)."""
    theta = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    unit = torch.stack((theta.cos(), theta.sin()), dim=-1)
    unit = unit / unit.abs().amax(dim=-1, keepdim=True)
    steps = torch.arange(1, n_points + 1, dtype=torch.float32)
    grid = unit[:, None, None, :] * steps[None, None, :, None]          # (heads, 1, points, 2)
    return grid.expand(n_heads, n_levels, n_points, 2).reshape(-1).clone()


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads:
            raise ValueError(
                "d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        head_dim = d_model // n_heads
        if head_dim & (head_dim - 1):
            # vector kernels need head_dim*sizeof(value) to be a power-of-two multiple of 16 B;
            # anything else still works through the generic kernels (the reference warns too)
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each "
                          "attention head a power of 2 which is more efficient in our CUDA "
                          "implementation.")
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.im2col_step = 128       # accepted by the Function for compatibility, ignored by the kernels
        self.fused_prologue = True   # softmax + location arithmetic inside the kernels when possible
        self.tensor_core_linears = True   # 3xTF32 tcgen05 GEMMs for the projections when eligible

        samples = n_heads * n_levels * n_points
        self.sampling_offsets = nn.Linear(d_model, samples * 2)
        self.attention_weights = nn.Linear(d_model, samples)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        linear_tc.install_invalidation_hooks(self)
        self._reset_parameters()

    def train(self, mode: bool = True):
        linear_tc.invalidate()   # weights may have been swapped (EMA, .data writes) between phases
        return super().train(mode)

    def _reset_parameters(self):
        linear_tc.invalidate()
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            self.sampling_offsets.bias = nn.Parameter(
                compass_offsets(self.n_heads, self.n_levels, self.n_points))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            for proj in (self.value_proj, self.output_proj):
                nn.init.xavier_uniform_(proj.weight)
                proj.bias.zero_()

    # -- pieces of forward ---------------------------------------------------------------------
    def _values(self, input_flatten, input_padding_mask):
        """(N, S, M, D) projected values with padded positions zeroed (mod.py:98-101)."""
        if self.tensor_core_linears:
            # the padding mask rides in the epilogue of the projection (PCT_LINEAR_ROWMASK)
            value = linear_tc.linear(input_flatten, self.value_proj.weight, self.value_proj.bias,
                                     row_mask=input_padding_mask)
        else:
            value = self.value_proj(input_flatten)
            if input_padding_mask is not None:
                value = value.masked_fill(input_padding_mask[..., None], 0.0)
        n, s, _ = value.shape
        return value.view(n, s, self.n_heads, self.d_model // self.n_heads)

    def _locations(self, reference_points, offsets, spatial_shapes):
        """Normalised sampling locations (N, Lq, M, L, P, 2) (mod.py:106-112)."""
        ref = reference_points[:, :, None, :, None, :]
        if reference_points.shape[-1] == 2:
            wh = spatial_shapes.flip(-1)                      # (L, 2) as (W, H)
            return ref + offsets / wh[None, None, None, :, None, :]
        return ref[..., :2] + offsets / self.n_points * ref[..., 2:] * 0.5

    def _check_lengths(self, spatial_shapes, len_in):
        """``sum_l H_l*W_l == Len_in`` (mod.py:96).  Reading the shapes back is a device sync, so a
        given shapes tensor is checked once, not in every layer of every step."""
        seen = self.__dict__.get("_lengths_ok")
        # keyed on the tensor OBJECT (weak reference), not on its address or version counter: an
        # address can be reused by another tensor, and inference tensors have no version counter
        if seen is not None and seen[0]() is spatial_shapes and seen[1] == len_in:
            return
        total = int(spatial_shapes.prod(1).sum())
        if total != len_in:
            raise ValueError(f"input_spatial_shapes describe {total} positions but input_flatten has "
                             f"{len_in} (mod.py:96)")
        self.__dict__["_lengths_ok"] = (weakref.ref(spatial_shapes), len_in)

    def _proj(self, layer, x):
        if self.tensor_core_linears:
            return linear_tc.linear(x, layer.weight, layer.bias)
        return layer(x)

    def _output(self, out, query, residual, norm):
        """output_proj, plus the caller's ``norm(residual + .)`` when it was handed in."""
        out = self._match(out, query)
        if norm is None:
            return self._proj(self.output_proj, out)
        if self.tensor_core_linears:
            return linear_tc.linear_residual_layernorm(out, self.output_proj.weight,
                                                       self.output_proj.bias, residual, norm)
        return norm(residual + self.output_proj(out))

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None, *, residual=None, norm=None):
        """query (N, Lq, C); reference_points (N, Lq, L, 2) in [0,1] or (N, Lq, L, 4) boxes
        (cx, cy, w, h); input_flatten (N, sum_l H_l*W_l, C); input_spatial_shapes (L, 2) rows
        (H_l, W_l); input_level_start_index (L,); input_padding_mask (N, S) bool, True = padding.
        Returns (N, Lq, C) -- or ``norm(residual + that)`` when the two keyword arguments (not in
        the reference's signature) are given."""
        n, len_q, _ = query.shape
        len_in = input_flatten.shape[1]
        self._check_lengths(input_spatial_shapes, len_in)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        m, lv, pt = self.n_heads, self.n_levels, self.n_points
        value = self._values(input_flatten, input_padding_mask)
        if self.tensor_core_linears:
            # both project the query: one GEMM with two dense results when the widths allow it
            offsets, logits = linear_tc.linear_pair(query, self.sampling_offsets, self.attention_weights)
        else:
            offsets, logits = self.sampling_offsets(query), self.attention_weights(query)
        offsets = offsets.view(n, len_q, m, lv, pt, 2)
        logits = logits.view(n, len_q, m, lv * pt)

        if self.fused_prologue and MSDeformAttnFusedFunction.supported(value, reference_points, lv, pt):
            out = MSDeformAttnFusedFunction.apply(
                value.contiguous(), input_spatial_shapes, input_level_start_index,
                reference_points.float().contiguous(), offsets.float().contiguous(),
                logits.float().contiguous())
            return self._output(out, query, residual, norm)

        weights = torch.softmax(logits, dim=-1).view(n, len_q, m, lv, pt)
        locations = self._locations(reference_points, offsets, input_spatial_shapes)
        if value.dtype in (torch.bfloat16, torch.float16):
            # low-precision values: bf16 storage, fp32 locations / weights, fp32 accumulation
            value = value.to(torch.bfloat16)
            locations, weights = locations.float(), weights.float()
        out = MSDeformAttnFunction.apply(
            value.contiguous(), input_spatial_shapes, input_level_start_index,
            locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self._output(out, query, residual, norm)

    @staticmethod
    def _match(out, query):
        """Kernel output back to the caller's dtype -- except under autocast, where the following
        Linear casts its input anyway and a round trip through fp32 would only cost two copies."""
        if out.dtype == query.dtype or (torch.is_autocast_enabled() and out.is_floating_point()):
            return out
        return out.to(query.dtype)
