"""fp32 ``Linear`` on the B200 tensor cores for the encoder layer around the op
(C ABI in include/pct_linear.h, kernel in csrc/linear_tc.cu).

The reference runs its pixel decoder in fp32 with autocast off (PD/msdeformattn.py:314-320), so
the projections of ``MSDeformAttn`` (OPS/modules/ms_deform_attn.py:55-58) and the FFN of the
encoder layer (PD/msdeformattn.py:107-120) are SIMT fp32 GEMMs.  ``linear`` computes the same
``x @ W^T + b`` with ``tcgen05.mma.kind::tf32`` and the 3xTF32 split (error ~1e-6 relative, i.e.
fp32-level, not TF32-level), with bias, ReLU and the residual add fused;
``linear_residual_layernorm`` also fuses the LayerNorm that follows ``output_proj`` / ``linear2``
(with autograd it additionally writes what LayerNorm's backward needs).

Operand precision (``PRECISION``): the default is the reference's fp32 bar (3xTF32).  Under
``torch.autocast("cuda", dtype=torch.bfloat16)`` -- or everywhere with ``PRECISION = "bf16"`` -- the
same kernels run with their operands rounded to bf16 (``PCT_LINEAR_BF16``: one ``kind::f16`` MMA
instead of three ``kind::tf32`` per K step, fp32 accumulation, fp32 tensors in memory, the same
fused epilogues), which is what autocast computes for a Linear except that the result is not
rounded to bf16; forward, input-gradient and weight-gradient GEMMs take the bf16 operands
(``BF16_WGRAD = False`` keeps the weight gradients on the 3xTF32 kernel).  The mode of a forward
call is remembered for its backward.

Scope: CUDA fp32 tensors with in/out features multiples of 16 (32 for the bf16 operands).  Anything
else (CPU tensors, bf16 tensors, fp16 autocast, odd sizes) is simply not this kernel's job and goes
to ``torch.nn.functional.linear`` exactly as the reference does; this is the caller's GEMM, not the
MSDeformAttn op, which never has a fallback.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _cabi

ENABLED = True            # process-wide switch (tests flip it to compare against torch)
PRECISION = "auto"        # "fp32": 3xTF32 always (not eligible under autocast); "bf16": bf16 operands
                          # always; "auto": bf16 operands under bf16 autocast, 3xTF32 otherwise
BF16_HIDDEN = True        # bf16-operand FFN blocks keep the hidden activation and its gradient in bf16
BF16_WGRAD = True         # bf16-operand calls also take their weight gradient with bf16 operands (as
                          # autocast does); False keeps the weight gradients on the 3xTF32 kernel
_EPOCH = 0                # bumped by invalidate(): every cached packed weight becomes stale


def invalidate() -> None:
    """Drop every cached packed weight (they are re-packed lazily on next use).

    The cache below notices in-place updates through autograd's version counter (optimizer
    steps, ``copy_`` under ``no_grad``) and reallocation -- but a write through ``.data``
    (``p.data.mul_()``, EMA swaps, hand-written checkpoint loaders) changes neither.  The modules
    of this package call ``invalidate()`` from ``_reset_parameters``, after ``load_state_dict`` and
    on every ``train()`` / ``eval()`` switch; code that writes weights through ``.data`` outside of
    those must call it too."""
    global _EPOCH
    _EPOCH += 1


def install_invalidation_hooks(module: torch.nn.Module) -> None:
    """``load_state_dict`` into ``module`` drops the cached packed weights."""
    module.register_load_state_dict_post_hook(lambda mod, incompatible_keys: invalidate())


def packed_weight(weight: torch.Tensor, transposed: bool = False, bf16: bool = False) -> torch.Tensor:
    """Packed operand of ``weight`` (or of ``weight.T`` for the input gradient) -- 3xTF32 pieces or,
    with ``bf16``, the bf16 rounding -- cached until the weight changes in place (optimizer step),
    is reallocated, or ``invalidate()`` runs."""
    slot = weight.__dict__.setdefault("_pct_packed", {})
    tag = (weight._version if not weight.is_inference() else -1, weight.data_ptr(), weight.device, _EPOCH)
    hit = slot.get((transposed, bf16))
    if hit is not None and hit[0] == tag:
        return hit[1]
    w = weight.detach()
    packed = _cabi.linear_pack_weight(w.t().contiguous() if transposed else w, bf16=bf16)
    slot[(transposed, bf16)] = (tag, packed)
    return packed


def bf16_operands() -> bool:
    """Whether a tensor-core Linear called now rounds its operands to bf16 (see ``PRECISION``)."""
    if PRECISION == "bf16":
        return True
    return (PRECISION == "auto" and torch.is_autocast_enabled("cuda")
            and torch.get_autocast_dtype("cuda") == torch.bfloat16)


def eligible(x: torch.Tensor, weight: torch.Tensor) -> bool:
    if not (ENABLED and x.is_cuda and weight.is_cuda and x.dtype == torch.float32
            and weight.dtype == torch.float32
            and weight.shape[0] % 16 == 0 and weight.shape[1] % 16 == 0 and x.numel() > 0
            and x.data_ptr() % 16 == 0 and x.storage_offset() % 4 == 0):
        return False
    if bf16_operands():
        # both GEMM directions run with bf16 operands: K of the forward and of the input gradient
        return weight.shape[0] % 32 == 0 and weight.shape[1] % 32 == 0
    return not torch.is_autocast_enabled("cuda")


def _tag(t: torch.Tensor):
    return (t._version if not t.is_inference() else -1, t.data_ptr())


def packed_pair(w_a, w_b, b_a, b_b, transposed: bool, bf16: bool):
    """Packed operand of the row-concatenated weight ``[w_a; w_b]`` (or of its transpose) and the
    concatenated bias, cached like ``packed_weight`` (on ``w_a``, keyed on all four tensors)."""
    slot = w_a.__dict__.setdefault("_pct_packed_pair", {})
    tag = (_tag(w_a), _tag(w_b), _tag(b_a), _tag(b_b), id(w_b), w_a.device, _EPOCH)
    hit = slot.get((transposed, bf16))
    if hit is not None and hit[0] == tag:
        return hit[1], hit[2]
    w = torch.cat((w_a.detach(), w_b.detach()), 0)
    packed = _cabi.linear_pack_weight(w.t().contiguous() if transposed else w, bf16=bf16)
    bias = torch.cat((b_a.detach(), b_b.detach())) if not transposed else None
    slot[(transposed, bf16)] = (tag, packed, bias)
    return packed, bias


def pair_eligible(x, lin_a: torch.nn.Linear, lin_b: torch.nn.Linear) -> bool:
    """Two Linears on the same input can run as one GEMM (``linear_pair``): both eligible, biases
    present, the first width a multiple of 32 (output split) and both widths fit the K step of the
    input-gradient GEMM; not under deterministic algorithms (the pair's weight gradient is split-K)."""
    if lin_a.bias is None or lin_b.bias is None or torch.are_deterministic_algorithms_enabled():
        return False
    if not (eligible(x, lin_a.weight) and eligible(x, lin_b.weight)):
        return False
    na, nb = lin_a.weight.shape[0], lin_b.weight.shape[0]
    kstep = 32 if bf16_operands() else 16
    return na % 32 == 0 and na % kstep == 0 and (na + nb) % kstep == 0


class LinearPairFunction(Function):
    """``(x @ W_a^T + b_a, x @ W_b^T + b_b)`` as one GEMM with two dense results; backward: one
    input-gradient GEMM over the pair of output gradients, one weight-gradient GEMM.  For
    ``sampling_offsets`` / ``attention_weights`` of MSDeformAttn, which both project the query
    (OPS/modules/ms_deform_attn.py:102-104): three kernels and an add fewer per layer and step."""

    @staticmethod
    def forward(ctx, x, w_a, b_a, w_b, b_b, bf16=False):
        na, nb, k = w_a.shape[0], w_b.shape[0], w_a.shape[1]
        x2 = x.reshape(-1, k).contiguous()
        packed, bias = packed_pair(w_a, w_b, b_a, b_b, False, bf16)
        ya, yb = _cabi.linear_forward_pair(x2, packed, na, nb, bias, bf16=bf16)
        ctx.bf16 = bf16
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, w_a, w_b, b_a, b_b)
        return ya.view(*x.shape[:-1], na), yb.view(*x.shape[:-1], nb)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_a, grad_b):
        x2, w_a, w_b, b_a, b_b = ctx.saved_tensors
        na, nb, k = w_a.shape[0], w_b.shape[0], w_a.shape[1]
        ga = grad_a.reshape(-1, na).contiguous()
        gb = grad_b.reshape(-1, nb).contiguous()
        gx = gwa = gwb = gba = gbb = None
        if ctx.needs_input_grad[0]:
            packed_t, _ = packed_pair(w_a, w_b, b_a, b_b, True, ctx.bf16)
            gx = _cabi.linear_forward_cat(ga, gb, packed_t, k, bf16=ctx.bf16).view(ctx.x_shape)
        if any(ctx.needs_input_grad[1:5]):
            gw, gbias = _cabi.linear_wgrad_pair(ga, gb, x2, bf16=ctx.bf16 and BF16_WGRAD)
            gwa, gwb, gba, gbb = gw[:na], gw[na:], gbias[:na], gbias[na:]
        return gx, gwa, gba, gwb, gbb, None


def linear_pair(x, lin_a: torch.nn.Linear, lin_b: torch.nn.Linear):
    """``(lin_a(x), lin_b(x))`` -- one tensor-core GEMM when ``pair_eligible``, two otherwise."""
    if pair_eligible(x, lin_a, lin_b):
        return LinearPairFunction.apply(x, lin_a.weight, lin_a.bias, lin_b.weight, lin_b.bias,
                                        bf16_operands())
    return linear(x, lin_a.weight, lin_a.bias), linear(x, lin_b.weight, lin_b.bias)


def _gemm(x2, weight, n, bf16, transposed=False, **epilogue):
    """``epilogue(x2 @ weight^T)`` (or ``x2 @ weight`` with ``transposed``: the input gradient)."""
    return _cabi.linear_forward(x2, packed_weight(weight, transposed, bf16), n, bf16=bf16, **epilogue)


class TensorCoreLinearFunction(Function):
    """y = [relu](x @ W^T + b [+ residual]); gradients for x, W, b and residual.  ``row_mask``
    (bool, one entry per row of x): masked rows of y are zeros (mod.py:100-101) -- written by the
    GEMM's epilogue -- and pass no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu, row_mask=None, bf16=False):
        n, k = weight.shape
        x2 = x.reshape(-1, k).contiguous()
        res2 = residual.reshape(-1, n).contiguous() if residual is not None else None
        b = bias.detach().contiguous() if bias is not None else None
        mask = row_mask.reshape(-1).contiguous() if row_mask is not None else None
        y = _gemm(x2, weight, n, bf16, bias=b, residual=res2, relu=relu, row_mask=mask)
        ctx.bf16 = bf16
        ctx.mask = mask
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, weight, y if relu else None)
        return y.view(*x.shape[:-1], n)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x2, weight, y = ctx.saved_tensors
        n, k = weight.shape
        g = grad_y.reshape(-1, n).contiguous()
        if ctx.relu:
            g = torch.ops.aten.threshold_backward(g, y, 0.0)     # one pass: g where y > 0 else 0
        if ctx.mask is not None:
            g = g.masked_fill(ctx.mask.view(-1, 1).bool(), 0.0)   # masked rows pass no gradient
        gx = gw = gb = gres = None
        if ctx.needs_input_grad[0]:
            # dX = dY @ W: the same kernel with the transposed weight as its operand
            gx = _gemm(g, weight, k, ctx.bf16, transposed=True).view(ctx.x_shape)
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            # dW = dY^T @ X split over the rows; the bias gradient rides on the same pass over dY
            gw, gb = _wgrad(g, x2, want_b, ctx.bf16)
        elif want_b:
            gb = g.sum(0)
        if ctx.has_res and ctx.needs_input_grad[3]:
            gres = g.view(*ctx.x_shape[:-1], n)
        return gx, gw, gb, gres, None, None, None


class LinearResidualLayerNormFunction(Function):
    """y = LayerNorm(x @ W^T + b + residual) in one kernel, with autograd: the kernel also writes
    LayerNorm's input and row statistics, the backward is the one-pass LayerNorm backward of
    csrc/layernorm_bwd.cu followed by the tensor-core input / weight gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, ln_weight, ln_bias, eps, bf16=False):
        n, k = weight.shape
        x2 = x.reshape(-1, k).contiguous()
        res2 = residual.reshape(-1, n).contiguous()
        b = bias.detach().contiguous() if bias is not None else None
        y, pre, mean, rstd = _gemm(
            x2, weight, n, bf16, bias=b, residual=res2, ln_weight=ln_weight.detach(),
            ln_bias=ln_bias.detach(), ln_eps=eps, save_for_backward=True)
        ctx.bf16 = bf16
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, weight, pre, mean, rstd, ln_weight, ln_bias)
        return y.view(*x.shape[:-1], n)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x2, weight, pre, mean, rstd, ln_weight, ln_bias = ctx.saved_tensors
        n, k = weight.shape
        gy = grad_y.reshape(-1, n).contiguous()
        g, g_lnw, g_lnb = _layernorm_backward(gy, pre, mean, rstd, ln_weight, ln_bias)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _gemm(g, weight, k, ctx.bf16, transposed=True).view(ctx.x_shape)
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            gw, gb = _wgrad(g, x2, want_b, ctx.bf16)
        elif want_b:
            gb = g.sum(0)
        gres = g.view(*ctx.x_shape[:-1], n) if ctx.needs_input_grad[3] else None
        return gx, gw, gb, gres, g_lnw, g_lnb, None, None


def _layernorm_backward(gy, pre, mean, rstd, ln_weight, ln_bias):
    """(d pre_norm, d gamma, d beta): the one-pass kernel of csrc/layernorm_bwd.cu -- or aten's under
    ``torch.use_deterministic_algorithms(True)`` (its column sums are fp32 reductions in arbitrary
    order) and for widths it does not cover."""
    n = gy.shape[1]
    if _cabi.layernorm_backward_supported(n) and not torch.are_deterministic_algorithms_enabled():
        return _cabi.layernorm_backward(gy, pre, mean, rstd, ln_weight.detach().contiguous())
    return torch.ops.aten.native_layer_norm_backward(gy, pre, [n], mean, rstd, ln_weight, ln_bias,
                                                     [True, True, True])


def _wgrad(g, x2, want_bias, bf16=False):
    """(dW, db): split-K tensor-core kernel (fp32 reductions, order varies run to run) -- or, under
    ``torch.use_deterministic_algorithms(True)``, torch's deterministic GEMM and column sum.
    ``bf16``: the forward ran with bf16 operands; so does the weight gradient unless
    ``BF16_WGRAD`` is off."""
    bf16 = bf16 and BF16_WGRAD and not torch.are_deterministic_algorithms_enabled()
    if not bf16:   # the 3xTF32 kernel and torch read fp32 (a bf16 hidden activation is widened)
        g, x2 = g.float(), x2.float()
    if torch.are_deterministic_algorithms_enabled():
        return g.t() @ x2, (g.sum(0) if want_bias else None)
    if want_bias:
        return _cabi.linear_wgrad(g, x2, with_bias=True, bf16=bf16)
    return _cabi.linear_wgrad(g, x2, bf16=bf16), None


class FFNFunction(Function):
    """y = LayerNorm(src + linear2(relu(linear1(src)))) -- the encoder layer's FFN block
    (PD/msdeformattn.py:107-120) as three kernels forward (GEMM+ReLU, GEMM+residual+LayerNorm) and
    five backward: the one-pass LayerNorm backward, two weight-gradient GEMMs (bias gradients included)
    and two input-gradient GEMMs whose epilogues carry the ReLU mask and the residual branch."""

    @staticmethod
    def forward(ctx, src, w1, b1, w2, b2, ln_weight, ln_bias, eps, bf16=False):
        c, f = w1.shape[1], w1.shape[0]
        x2 = src.reshape(-1, c).contiguous()
        # bf16 operands: the hidden activation is only ever a GEMM operand or a gate, both of which
        # round it to bf16 anyway -- it is stored as bf16 (half the traffic of the block's largest
        # tensor, no result changes); likewise its gradient in the backward
        low = torch.bfloat16 if (bf16 and BF16_HIDDEN) else None
        h = _gemm(x2, w1, f, bf16, bias=b1.detach(), relu=True, out_dtype=low)
        y, pre, mean, rstd = _gemm(
            h, w2, c, bf16, bias=b2.detach(), residual=x2, ln_weight=ln_weight.detach(),
            ln_bias=ln_bias.detach(), ln_eps=eps, save_for_backward=True)
        ctx.bf16 = bf16
        ctx.src_shape = src.shape
        ctx.save_for_backward(x2, w1, w2, h, pre, mean, rstd, ln_weight, ln_bias)
        return y.view(src.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x2, w1, w2, h, pre, mean, rstd, ln_weight, ln_bias = ctx.saved_tensors
        c, f = w1.shape[1], w1.shape[0]
        gy = grad_y.reshape(-1, c).contiguous()
        g, g_lnw, g_lnb = _layernorm_backward(gy, pre, mean, rstd, ln_weight, ln_bias)
        gw2, gb2 = _wgrad(g, h, True, ctx.bf16)
        # d(hidden) = (g @ W2) masked by the ReLU, straight out of the GEMM's epilogue
        gh = _gemm(g, w2, f, ctx.bf16, transposed=True, gate=h, out_dtype=h.dtype)
        gw1, gb1 = _wgrad(gh, x2, True, ctx.bf16)
        # d(src) = gh @ W1 + g (the residual branch), the add fused as well
        gx = _gemm(gh, w1, c, ctx.bf16, transposed=True, residual=g)
        return gx.view(ctx.src_shape), gw1, gb1, gw2, gb2, g_lnw, g_lnb, None, None


def ffn_block(src, linear1, linear2, norm):
    """``norm(src + linear2(relu(linear1(src))))`` with no dropout in between; picks the fused
    autograd Function, the fused inference kernels or the plain composition."""
    ok = (eligible(src, linear1.weight) and eligible(src, linear2.weight)
          and linear1.bias is not None and linear2.bias is not None and norm.elementwise_affine
          and norm.bias is not None and linear2.weight.shape[0] <= 256
          and tuple(norm.normalized_shape) == (linear2.weight.shape[0],))
    needs_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in (src, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
                                  norm.weight, norm.bias))
    if ok and needs_grad:
        return FFNFunction.apply(src, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
                                 norm.weight, norm.bias, norm.eps, bf16_operands())
    if ok and bf16_operands() and BF16_HIDDEN:
        # inference with bf16 operands: the same two kernels with the hidden activation kept in bf16
        c, f = linear1.weight.shape[1], linear1.weight.shape[0]
        x2 = src.reshape(-1, c).contiguous()
        h = _gemm(x2, linear1.weight, f, True, bias=linear1.bias.detach(), relu=True,
                  out_dtype=torch.bfloat16)
        y = _gemm(h, linear2.weight, c, True, bias=linear2.bias.detach(), residual=x2,
                  ln_weight=norm.weight.detach(), ln_bias=norm.bias.detach(), ln_eps=norm.eps)
        return y.view(src.shape)
    hidden = linear(src, linear1.weight, linear1.bias, relu=True)
    return linear_residual_layernorm(hidden, linear2.weight, linear2.bias, src, norm)


def linear(x, weight, bias=None, residual=None, relu: bool = False, row_mask=None):
    """``[relu](F.linear(x, weight, bias) [+ residual])`` -- on the tensor cores when eligible.
    ``row_mask`` (bool, shape of x without its last dim): ``.masked_fill(row_mask[..., None], 0)``
    on the result, done by the GEMM's epilogue on the tensor-core path."""
    if eligible(x, weight) and (row_mask is None or (residual is None and row_mask.is_cuda)):
        return TensorCoreLinearFunction.apply(x, weight, bias, residual, relu, row_mask,
                                              bf16_operands())
    y = F.linear(x, weight, bias)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return y.masked_fill(row_mask[..., None], 0.0) if row_mask is not None else y


def linear_residual_layernorm(x, weight, bias, residual, norm: torch.nn.LayerNorm):
    """``norm(residual + F.linear(x, weight, bias))`` as one kernel.  With autograd the kernel also
    writes LayerNorm's input and row statistics for the backward pass."""
    n = weight.shape[0]
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x, weight, bias, residual, norm.weight, norm.bias))
    fusable = (eligible(x, weight) and n <= 256 and norm.elementwise_affine
               and norm.bias is not None and tuple(norm.normalized_shape) == (n,)
               and norm.weight.dtype == torch.float32)
    if not fusable:
        return norm(linear(x, weight, bias, residual=residual))
    if needs_grad:
        return LinearResidualLayerNormFunction.apply(x, weight, bias, residual, norm.weight,
                                                     norm.bias, norm.eps, bf16_operands())
    x2 = x.reshape(-1, weight.shape[1]).contiguous()
    y = _gemm(x2, weight, n, bf16_operands(), bias=bias, residual=residual.reshape(-1, n).contiguous(),
              ln_weight=norm.weight, ln_bias=norm.bias, ln_eps=norm.eps)
    return y.view(*x.shape[:-1], n)
