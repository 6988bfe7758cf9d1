"""``MSDeformAttnFunction`` -- drop-in for OPS/functions/ms_deform_attn_func.py:32-49.

Same positional signature and return convention as the reference::

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step) -> output

    value               (N, S, M, D)          float32 | bfloat16 | float64, CUDA, contiguous
    value_spatial_shapes (L, 2) int64 CUDA    rows (H_l, W_l)
    value_level_start_index (L,) int64 CUDA
    sampling_locations  (N, Lq, M, L, P, 2)   (x, y) normalised to [0, 1]
    attention_weights   (N, Lq, M, L, P)
    output              (N, Lq, M*D)

``backward`` returns ``(grad_value, None, None, grad_sampling_loc, grad_attn_weight, None)``
(func.py:49).  The work is done by the sm_100a kernels behind the C ABI in
``include/pct_msda.h``; the preconditions the reference asserts in C++
(OPS/src/cuda/ms_deform_attn_cuda.cu:33-57) are checked here and raise ``RuntimeError`` /
``ValueError``.  ``im2col_step`` is accepted for compatibility and ignored: the whole batch goes
through one launch, so the reference's ``batch % im2col_step == 0`` restriction (cu:57) is gone.

bf16: ``value`` may be bfloat16 (not supported by the reference, cu:69).  Locations and weights
are then taken in float32 -- what type promotion produces in mod.py:104-109 under autocast --
accumulation is fp32, output / grad_value are bf16.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

# Tuning hook for A/B measurements (scripts/): option bits of include/pct_msda.h added to every
# backward call, e.g. PCT_MSDA_BWD_FLAGS=0x20 forces the per-tap gradient kernel.  Default 0.
import os as _os
_TUNE_BWD_FLAGS = int(_os.environ.get("PCT_MSDA_BWD_FLAGS", "0"), 0)

from .. import _cabi


def _require(cond: bool, msg: str, exc=RuntimeError):
    if not cond:
        raise exc(msg)


def _check_inputs(value, shapes, lsi, loc, attn):
    for name, t in (("value", value), ("value_spatial_shapes", shapes),
                    ("value_level_start_index", lsi), ("sampling_locations", loc),
                    ("attention_weights", attn)):
        _require(isinstance(t, torch.Tensor), f"{name} must be a tensor", TypeError)
        _require(t.is_cuda, f"{name} must be a CUDA tensor")            # cu:39-43
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")   # cu:33-37
        _require(t.device == value.device, f"{name} must be on the same device as value")
    _require(value.dtype in _cabi.DTYPE_CODES,
             f"value dtype {value.dtype} not supported (float32, bfloat16, float64)", TypeError)
    loc_dtype = torch.float64 if value.dtype == torch.float64 else torch.float32
    _require(loc.dtype == loc_dtype and attn.dtype == loc_dtype,
             f"sampling_locations / attention_weights must be {loc_dtype} for {value.dtype} value "
             f"(got {loc.dtype}, {attn.dtype})", TypeError)
    _require(shapes.dtype == torch.int64 and lsi.dtype == torch.int64,
             "value_spatial_shapes / value_level_start_index must be int64", TypeError)   # cu:72-73
    _require(value.dim() == 4, "value must be (N, S, M, D)", ValueError)
    _require(shapes.dim() == 2 and shapes.shape[1] == 2, "value_spatial_shapes must be (L, 2)", ValueError)
    N, S, M, D = value.shape
    L = shapes.shape[0]
    _require(lsi.shape == (L,), "value_level_start_index must be (L,)", ValueError)
    _require(loc.dim() == 6 and loc.shape[0] == N and loc.shape[2] == M and loc.shape[3] == L
             and loc.shape[5] == 2, "sampling_locations must be (N, Lq, M, L, P, 2)", ValueError)
    _require(tuple(attn.shape) == tuple(loc.shape[:5]),
             "attention_weights must be (N, Lq, M, L, P)", ValueError)


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        _check_inputs(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                      attention_weights)
        output = _cabi.forward(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        # the reference asserts contiguity (cu:103); autograd may hand us an expanded or
        # transposed grad, so make it dense instead of failing
        grad_output = grad_output.contiguous()
        if grad_output.dtype != value.dtype:
            grad_output = grad_output.to(value.dtype)
        flags = 0
        if torch.are_deterministic_algorithms_enabled():
            # bit-reproducible grad_value (fixed-point integer reductions); the reference's float
            # atomics (cuh:130-157) have no such mode
            N, S, M, D = value.shape
            supported = (value.dtype in (torch.float32, torch.bfloat16) and value.numel() > 0
                         and _cabi.kernel_family(D, shapes.shape[0], loc.shape[4], value.dtype, True) == 1)
            if supported:
                flags |= _cabi.FLAG_DETERMINISTIC
            elif not torch.is_deterministic_algorithms_warn_only_enabled():
                raise RuntimeError("MSDeformAttnFunction.backward has no deterministic implementation "
                                   f"for value dtype {value.dtype} / head dim {D}")
        try:
            grad_value, grad_loc, grad_attn = _cabi.backward(value, shapes, lsi, loc, attn,
                                                             grad_output, flags=flags or _TUNE_BWD_FLAGS,
                                                             host_shapes=_cabi.host_shapes_of(shapes))
        except _cabi.MSDAError as e:
            if not (flags and e.code == _cabi.E_UNSUPPORTED):
                raise
            # vector-family shape without a deterministic kernel (head rows wider than 512 bytes)
            if not torch.is_deterministic_algorithms_warn_only_enabled():
                raise RuntimeError("MSDeformAttnFunction.backward has no deterministic implementation "
                                   f"for value dtype {value.dtype} / head dim {value.shape[3]}") from e
            grad_value, grad_loc, grad_attn = _cabi.backward(value, shapes, lsi, loc, attn, grad_output)
        return grad_value, None, None, grad_loc, grad_attn, None


class MSDeformAttnFusedFunction(Function):
    """Fused-prologue variant: takes what the module's Linears produce and
    evaluates the softmax over L*P (mod.py:104) and the location arithmetic (mod.py:106-112)
    inside the kernels, so neither ``sampling_locations`` nor ``attention_weights`` (nor their
    gradients) ever exist in HBM.

        MSDeformAttnFusedFunction.apply(value, value_spatial_shapes, value_level_start_index,
                                        reference_points, sampling_offsets, attention_logits)

    reference_points (N, Lq, L, 2|4) fp32, no gradient; sampling_offsets (N, Lq, M, L, P, 2) and
    attention_logits (N, Lq, M, L*P) fp32, raw.  ``supported(...)`` tells whether a call can take
    this route; otherwise the caller composes the prologue in PyTorch and uses
    ``MSDeformAttnFunction`` (still the CUDA kernels -- there is no CPU route either way)."""

    @staticmethod
    def supported(value, reference_points, n_levels, n_points):
        return (value.is_cuda and value.dtype in (torch.float32, torch.bfloat16)
                and not torch.are_deterministic_algorithms_enabled()   # deterministic mode: unfused route
                and not reference_points.requires_grad and value.shape[1] > 0
                and reference_points.shape[-1] in (2, 4)
                and _cabi.kernel_family(value.shape[3], n_levels, n_points, value.dtype) == 1
                and _cabi.kernel_family(value.shape[3], n_levels, n_points, value.dtype, True) == 1)

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points,
                sampling_offsets, attention_logits):
        N, S, M, D = value.shape
        L = value_spatial_shapes.shape[0]
        _check_inputs(value, value_spatial_shapes, value_level_start_index, sampling_offsets,
                      attention_logits.view(tuple(sampling_offsets.shape[:5])))
        _require(reference_points.is_cuda and reference_points.is_contiguous()
                 and reference_points.dtype == torch.float32,
                 "reference_points must be a contiguous float32 CUDA tensor")
        _require(tuple(reference_points.shape[:3]) == (N, sampling_offsets.shape[1], L)
                 and reference_points.shape[-1] in (2, 4),
                 "reference_points must be (N, Lq, L, 2) or (N, Lq, L, 4)", ValueError)
        out = _cabi.fused_forward(value, value_spatial_shapes, value_level_start_index,
                                  reference_points, sampling_offsets, attention_logits)
        _require(out is not None, "shape not supported by the fused MSDeformAttn kernels; "
                                  "check MSDeformAttnFusedFunction.supported() first")
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              reference_points, sampling_offsets, attention_logits)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, ref, offsets, logits = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        if grad_output.dtype != value.dtype:
            grad_output = grad_output.to(value.dtype)
        gv, goff, glog = _cabi.fused_backward(value, shapes, lsi, ref, offsets, logits, grad_output,
                                              flags=_TUNE_BWD_FLAGS,
                                              host_shapes=_cabi.host_shapes_of(shapes))
        return gv, None, None, None, goff, glog


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Debug / test helper kept because callers import this name from this module
    (OPS/modules/ms_deform_attn.py:25, OPS/test.py:21; func.py:53-54 says "for debug and test
    only").  Plain, differentiable PyTorch written as explicit index gathers of the four bilinear
    corners (not ``grid_sample``), so it doubles as an independent cross-check of the kernels.
    NEVER called by MSDeformAttnFunction or the MSDeformAttn module: the product path has no
    fallback."""
    N, S, M, D = value.shape
    Lq, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    by_head = value.permute(0, 2, 1, 3)                                   # (N, M, S, D)
    out = value.new_zeros((N, Lq, M, D))
    start = 0
    for lvl, (H, W) in enumerate(value_spatial_shapes.tolist()):
        if H * W == 0:
            continue
        x = sampling_locations[:, :, :, lvl, :, 0] * W - 0.5              # (N, Lq, M, P)
        y = sampling_locations[:, :, :, lvl, :, 1] * H - 0.5
        x0, y0 = x.floor(), y.floor()
        lx, ly = x - x0, y - y0
        a = attention_weights[:, :, :, lvl]
        corners = ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx),
                   (1, 0, ly * (1 - lx)), (1, 1, ly * lx))
        for dy, dx, w in corners:
            yi, xi = y0.long() + dy, x0.long() + dx
            ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
            flat = start + yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)    # (N, Lq, M, P)
            idx = flat.permute(0, 2, 1, 3).reshape(N, M, Lq * P, 1).expand(-1, -1, -1, D)
            taps = by_head.gather(2, idx).view(N, M, Lq, P, D).permute(0, 2, 1, 3, 4)
            out = out + (taps * (w * a * ok)[..., None]).sum(3)
        start += H * W
    return out.reshape(N, Lq, M * D)
