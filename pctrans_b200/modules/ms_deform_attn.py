"""``MSDeformAttn`` module -- drop-in for OPS/modules/ms_deform_attn.py:34-125.

Same constructor ``(d_model=256, n_levels=4, n_heads=8, n_points=4)``, same parameter names
(``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` -- so PCTrans
checkpoints load under ``...encoder.layers.{i}.self_attn.*``), same initialisation and the same
``forward(query, reference_points, input_flatten, input_spatial_shapes,
input_level_start_index, input_padding_mask=None)``.

Differences, all deliberate:
  * the reference wraps the op in a bare ``try/except`` and silently falls back to its PyTorch
    path on ANY error (mod.py:116-121); here errors propagate and there is no fallback;
  * ``im2col_step`` is kept as an attribute (mod.py:52) but the kernels ignore it.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(
                "d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        head_dim = d_model // n_heads
        if not _is_power_of_2(head_dim):
            # the 128-bit vector kernels need head_dim * sizeof(value) to be a power-of-two
            # multiple of 16 bytes; other sizes run the generic kernels (mod.py:46-50 warns too)
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each "
                          "attention head a power of 2 which is more efficient in our CUDA "
                          "implementation.")

        self.im2col_step = 128
        # Use the fused-prologue kernels (softmax + location arithmetic inside the op, SURVEY.md
        # 8f rank 1) when the call qualifies; results agree with the unfused route to rounding.
        self.fused_prologue = True
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """mod.py:64-80: zero offset weights, compass-direction bias scaled by the point index,
        zero attention logits, Xavier projections."""
        constant_(self.sampling_offsets.weight.data, 0.0)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        bias = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for k in range(self.n_points):
            bias[:, :, k, :] *= k + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.view(-1))
        constant_(self.attention_weights.weight.data, 0.0)
        constant_(self.attention_weights.bias.data, 0.0)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.0)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None):
        """
        query                   (N, Lq, C)
        reference_points        (N, Lq, L, 2) in [0,1] (top-left (0,0), bottom-right (1,1),
                                including padding) or (N, Lq, L, 4) boxes (cx, cy, w, h)
        input_flatten           (N, sum_l H_l*W_l, C)
        input_spatial_shapes    (L, 2)  [(H_0, W_0), ...]
        input_level_start_index (L,)
        input_padding_mask      (N, sum_l H_l*W_l) bool, True on padding
        returns                 (N, Lq, C)
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        offsets = self.sampling_offsets(query).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(
            N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        if self.fused_prologue and MSDeformAttnFusedFunction.supported(
                value, reference_points, self.n_levels, self.n_points):
            output = MSDeformAttnFusedFunction.apply(
                value.contiguous(), input_spatial_shapes, input_level_start_index,
                reference_points.float().contiguous(), offsets.float().contiguous(),
                weights.float().contiguous())
            return self.output_proj(output.to(query.dtype))
        weights = F.softmax(weights, -1).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack(
                [input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = reference_points[:, :, None, :, None, :] \
                + offsets / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = reference_points[:, :, None, :, None, :2] \
                + offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        # no try/except and no PyTorch fallback: a failure of the CUDA path is an error
        if value.dtype == torch.bfloat16 or value.dtype == torch.float16:
            value = value.to(torch.bfloat16)
            locations, weights = locations.float(), weights.float()
        output = MSDeformAttnFunction.apply(
            value.contiguous(), input_spatial_shapes, input_level_start_index,
            locations.contiguous(), weights.contiguous(), self.im2col_step)
        output = self.output_proj(output.to(query.dtype))
        return output
