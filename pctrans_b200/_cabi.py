"""ctypes binding of the C ABI in ``include/pct_msda.h`` (``libpctrans_msda.so``).

This is the only route from Python to the kernels.  There is no fallback: if the shared library is
missing or a call fails, an exception is raised (the reference's module silently reroutes every
failure to its slow PyTorch path, OPS/modules/ms_deform_attn.py:119-121 -- deliberately not
reproduced).
"""
from __future__ import annotations

import ctypes
import weakref
import os
import threading

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# PCTRANS_MSDA_LIB selects another build of the SAME C ABI (the bounds-checking debug library in
# tests); it is not a fallback: a missing file is still an ImportError.
LIB_PATH = os.environ.get("PCTRANS_MSDA_LIB") or os.path.join(_PKG_DIR, "_build", "libpctrans_msda.so")
ABI_VERSION = 4

DTYPE_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float64: 2}

EXPORTS = (
    "pct_msda_abi_version", "pct_msda_error_string", "pct_msda_launch_count",
    "pct_msda_kernel_family", "pct_msda_backward_plan", "pct_msda_forward", "pct_msda_forward_ex",
    "pct_msda_backward_workspace_bytes", "pct_msda_backward_workspace_bytes_ex",
    "pct_msda_backward", "pct_msda_backward_ex",
    "pct_msda_fused_forward", "pct_msda_fused_backward", "pct_msda_fused_backward_ex",
    "pct_msda_fused_backward_ex2", "pct_msda_backward_ex2", "pct_msda_backward_plan_ex",
    "pct_msda_host_ctx_create", "pct_msda_host_ctx_destroy", "pct_msda_host_forward",
    "pct_msda_host_forward_backward",
)
LINEAR_EXPORTS = ("pct_linear_packed_bytes", "pct_linear_pack_weight", "pct_linear_packed_bytes_bf16",
                  "pct_linear_pack_weight_bf16", "pct_linear_forward",
                  "pct_linear_forward_train", "pct_linear_forward_split", "pct_linear_wgrad", "pct_linear_wgrad_ex",
                  "pct_linear_wgrad_split", "pct_linear_layernorm_backward")
LINEAR_RELU, LINEAR_RESIDUAL, LINEAR_LAYERNORM, LINEAR_GATE, LINEAR_ROWMASK = 0x1, 0x2, 0x4, 0x8, 0x10
LINEAR_BF16, LINEAR_X_BF16, LINEAR_Y_BF16, LINEAR_RES_BF16, LINEAR_GY_BF16 = 0x20, 0x40, 0x80, 0x100, 0x200
E_UNSUPPORTED = 1007
FLAG_FORCE_GENERIC = 0x1
FLAG_BF16_ACCUM = 0x8
FLAG_DETERMINISTIC = 0x10
FLAG_BWD_PER_TAP = 0x20
FLAG_BWD_BINNED = 0x40
FLAG_BF16_F16_ACCUM = 0x80

_lib = None
_lock = threading.Lock()


class MSDAError(RuntimeError):
    """A pct_msda_* entry point returned a non-zero code (``.code``)."""

    def __init__(self, msg, code=0):
        super().__init__(msg)
        self.code = code


def load():
    """Load (once) and return the ctypes handle; raises ImportError if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m pctrans_b200.build` "
                "(nvcc, sm_100a).  pctrans_b200 has no CPU or PyTorch fallback for MSDeformAttn.")
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, u32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t
        lib.pct_msda_abi_version.restype = i32
        lib.pct_msda_abi_version.argtypes = []
        lib.pct_msda_error_string.restype = ctypes.c_char_p
        lib.pct_msda_error_string.argtypes = [i32]
        lib.pct_msda_launch_count.restype = ctypes.c_uint64
        lib.pct_msda_launch_count.argtypes = []
        lib.pct_msda_kernel_family.restype = i32
        lib.pct_msda_kernel_family.argtypes = [i32] * 5
        lib.pct_msda_backward_plan.restype = i32
        lib.pct_msda_backward_plan.argtypes = [i32] * 8 + [u32]
        fwd_args = [vp] * 6 + [i32] * 8 + [vp]
        lib.pct_msda_forward.restype = i32
        lib.pct_msda_forward.argtypes = fwd_args
        lib.pct_msda_forward_ex.restype = i32
        lib.pct_msda_forward_ex.argtypes = fwd_args + [u32]
        lib.pct_msda_backward_workspace_bytes.restype = sz
        lib.pct_msda_backward_workspace_bytes.argtypes = [i32] * 5
        lib.pct_msda_backward_workspace_bytes_ex.restype = sz
        lib.pct_msda_backward_workspace_bytes_ex.argtypes = [i32] * 5 + [u32]
        bwd_args = [vp] * 9 + [i32] * 8 + [vp, sz, vp]
        lib.pct_msda_backward.restype = i32
        lib.pct_msda_backward.argtypes = bwd_args
        lib.pct_msda_backward_ex.restype = i32
        lib.pct_msda_backward_ex.argtypes = bwd_args + [u32]
        lib.pct_msda_fused_forward.restype = i32
        lib.pct_msda_fused_forward.argtypes = [vp] * 4 + [i32] + [vp] * 3 + [i32] * 8 + [vp]
        lib.pct_msda_fused_backward.restype = i32
        lib.pct_msda_fused_backward.argtypes = [vp] * 4 + [i32] + [vp] * 6 + [i32] * 8 + [vp, sz, vp]
        lib.pct_msda_fused_backward_ex.restype = i32
        lib.pct_msda_fused_backward_ex.argtypes = [vp] * 4 + [i32] + [vp] * 6 + [i32] * 8 + [vp, sz, vp, u32]
        lib.pct_msda_fused_backward_ex2.restype = i32
        lib.pct_msda_fused_backward_ex2.argtypes = [vp] * 4 + [i32] + [vp] * 6 + [i32] * 8 + [vp, sz, vp, u32, vp]
        lib.pct_msda_backward_ex2.restype = i32
        lib.pct_msda_backward_ex2.argtypes = bwd_args + [u32, vp]
        lib.pct_msda_backward_plan_ex.restype = i32
        lib.pct_msda_backward_plan_ex.argtypes = [i32] * 8 + [u32, vp]
        # include/pct_linear.h
        lib.pct_linear_packed_bytes.restype = sz
        lib.pct_linear_packed_bytes.argtypes = [i32, i32]
        lib.pct_linear_pack_weight.restype = i32
        lib.pct_linear_pack_weight.argtypes = [vp, vp, i32, i32, vp]
        lib.pct_linear_packed_bytes_bf16.restype = sz
        lib.pct_linear_packed_bytes_bf16.argtypes = [i32, i32]
        lib.pct_linear_pack_weight_bf16.restype = i32
        lib.pct_linear_pack_weight_bf16.argtypes = [vp, vp, i32, i32, vp]
        lib.pct_linear_wgrad.restype = i32
        lib.pct_linear_wgrad.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
        lib.pct_linear_wgrad_ex.restype = i32
        lib.pct_linear_wgrad_ex.argtypes = [vp, vp, vp, vp, i32, i32, i32, u32, vp]
        lib.pct_linear_wgrad_split.restype = i32
        lib.pct_linear_wgrad_split.argtypes = [vp, vp, i32, i32, vp, vp, vp, i32, i32, u32, vp]
        lib.pct_linear_forward_split.restype = i32
        lib.pct_linear_forward_split.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, u32, vp]
        lib.pct_linear_layernorm_backward.restype = i32
        lib.pct_linear_layernorm_backward.argtypes = [vp] * 8 + [i32, i32, vp]
        lib.pct_linear_forward_train.restype = i32
        lib.pct_linear_forward_train.argtypes = [vp] * 6 + [ctypes.c_float] + [vp] * 4 + [i32] * 3 + [u32, vp]
        lib.pct_linear_forward.restype = i32
        lib.pct_linear_forward.argtypes = [vp] * 6 + [ctypes.c_float, vp] + [i32] * 3 + [u32, vp]
        got = lib.pct_msda_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().pct_msda_error_string(code).decode()
        raise MSDAError(f"{what} failed with code {code}: {msg}", code)


def launch_count() -> int:
    return int(load().pct_msda_launch_count())


def kernel_family(channels: int, num_levels: int, num_point: int, dtype: torch.dtype,
                  backward: bool = False) -> int:
    return int(load().pct_msda_kernel_family(channels, num_levels, num_point, DTYPE_CODES[dtype],
                                             int(backward)))


# Host copies of spatial_shapes tensors, so that the backward can tell the library the exact cell
# count of the pyramid (pct_msda_backward_ex2).  Keyed on the tensor OBJECT; a tensor is read back
# once (one device sync per pyramid, like the reference's own length assert, mod.py:96) and never
# while a CUDA graph is being captured.
_HOST_SHAPES = {}


def register_host_shapes(spatial_shapes: torch.Tensor, shapes) -> None:
    """Tell the binding what a device ``spatial_shapes`` tensor holds (callers that built it from
    Python numbers: no device sync at all)."""
    arr = (ctypes.c_int64 * (2 * len(shapes)))(*[int(v) for hw in shapes for v in hw])
    key = id(spatial_shapes)
    _HOST_SHAPES[key] = (weakref.ref(spatial_shapes, lambda _r, k=key: _HOST_SHAPES.pop(k, None)),
                         getattr(spatial_shapes, "_version", 0) if not spatial_shapes.is_inference() else 0,
                         arr)


def host_shapes_of(spatial_shapes: torch.Tensor):
    """ctypes int64 array (L*2) with the rows of ``spatial_shapes``, or None when it cannot be had
    without disturbing the caller (graph capture in progress)."""
    hit = _HOST_SHAPES.get(id(spatial_shapes))
    if hit is not None and hit[0]() is spatial_shapes:
        ver = getattr(spatial_shapes, "_version", 0) if not spatial_shapes.is_inference() else 0
        if ver == hit[1]:
            return hit[2]
    if spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
        return None
    register_host_shapes(spatial_shapes, spatial_shapes.tolist())
    return _HOST_SHAPES[id(spatial_shapes)][2]


def backward_plan(N, S, M, D, L, Lq, P, dtype: torch.dtype, flags: int = 0, host_shapes=None) -> int:
    """0 generic, 1 per-tap vector scatter, 2 cell-binned (needs a current CUDA device).
    ``host_shapes``: what ``host_shapes_of`` returns, or a sequence of (H, W)."""
    if host_shapes is not None and not isinstance(host_shapes, ctypes.Array):
        host_shapes = (ctypes.c_int64 * (2 * len(host_shapes)))(*[int(v) for hw in host_shapes for v in hw])
    rc = int(load().pct_msda_backward_plan_ex(N, S, M, D, L, Lq, P, DTYPE_CODES[dtype], flags,
                                              host_shapes))
    if rc < 0:
        check(-rc, "pct_msda_backward_plan")
    return rc


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, flags: int = 0,
            out=None):
    """Raw forward on validated, contiguous CUDA tensors.  Returns output (N, Lq, M*D)
    (written into ``out`` when given)."""
    lib = load()
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if out is None:
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device_of(value):
        rc = lib.pct_msda_forward_ex(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(),
            N, S, M, D, L, Lq, P, DTYPE_CODES[value.dtype], _stream_ptr(value.device), flags)
    check(rc, "pct_msda_forward")
    return out


def backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
             flags: int = 0, out=None, workspace=None, host_shapes=None):
    """Raw backward.  Returns (grad_value, grad_sampling_loc, grad_attn_weight); ``out`` may
    supply the three tensors and ``workspace`` a uint8 scratch tensor to avoid allocations."""
    lib = load()
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    code = DTYPE_CODES[value.dtype]
    if out is None:
        grad_value = torch.empty_like(value)
        grad_loc = torch.empty_like(sampling_loc)
        grad_attn = torch.empty_like(attn_weight)
    else:
        grad_value, grad_loc, grad_attn = out
    ws_bytes = int(lib.pct_msda_backward_workspace_bytes_ex(N, S, M, D, code, flags))
    ws = workspace
    if ws_bytes and (ws is None or ws.numel() < ws_bytes):
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device)
    with torch.cuda.device_of(value):
        rc = lib.pct_msda_backward_ex2(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
            N, S, M, D, L, Lq, P, code, ws.data_ptr() if ws is not None else None, ws_bytes,
            _stream_ptr(value.device), flags, host_shapes)
    check(rc, "pct_msda_backward")
    return grad_value, grad_loc, grad_attn


def fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                  attn_logits, out=None):
    """Fused-prologue forward (softmax + location arithmetic inside the kernel).  Returns the
    output, or None when the shape is not served by the fused kernels (caller uses ``forward``)."""
    lib = load()
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
    if out is None:
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device_of(value):
        rc = lib.pct_msda_fused_forward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            reference_points.data_ptr(), reference_points.shape[-1], sampling_offsets.data_ptr(),
            attn_logits.data_ptr(), out.data_ptr(), N, S, M, D, L, Lq, P,
            DTYPE_CODES[value.dtype], _stream_ptr(value.device))
    if rc == E_UNSUPPORTED:
        return None
    check(rc, "pct_msda_fused_forward")
    return out


def fused_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                   attn_logits, grad_output, workspace=None, flags=0, host_shapes=None):
    """Returns (grad_value, grad_sampling_offsets, grad_attn_logits)."""
    lib = load()
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
    code = DTYPE_CODES[value.dtype]
    grad_value = torch.empty_like(value)
    grad_off = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attn_logits)
    ws_bytes = int(lib.pct_msda_backward_workspace_bytes(N, S, M, D, code))
    ws = workspace
    if ws_bytes and (ws is None or ws.numel() < ws_bytes):
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device)
    with torch.cuda.device_of(value):
        rc = lib.pct_msda_fused_backward_ex2(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            reference_points.data_ptr(), reference_points.shape[-1], sampling_offsets.data_ptr(),
            attn_logits.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(),
            grad_off.data_ptr(), grad_logits.data_ptr(), N, S, M, D, L, Lq, P, code,
            ws.data_ptr() if ws is not None else None, ws_bytes, _stream_ptr(value.device), flags,
            host_shapes)
    check(rc, "pct_msda_fused_backward_ex2")
    return grad_value, grad_off, grad_logits


def linear_pack_weight(weight, bf16: bool = False):
    """(out, in) fp32 weight -> packed operand (uint8 tensor) for linear_forward: the 3xTF32 split,
    or with ``bf16`` the bf16 rounding (for ``linear_forward(..., bf16=True)``)."""
    w = weight.detach().contiguous()
    n, k = w.shape
    lib = load()
    size_fn, pack_fn = ((lib.pct_linear_packed_bytes_bf16, lib.pct_linear_pack_weight_bf16) if bf16
                        else (lib.pct_linear_packed_bytes, lib.pct_linear_pack_weight))
    nbytes = size_fn(k, n)
    if nbytes == 0:
        raise MSDAError(f"pct_linear: unsupported Linear dimensions in={k} out={n}"
                        + (" (bf16 operands)" if bf16 else ""), 1002)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    with torch.cuda.device_of(w):
        check(pack_fn(w.data_ptr(), packed.data_ptr(), k, n, _stream_ptr(w.device)),
              "pct_linear_pack_weight")
    return packed


def linear_forward(x, w_packed, out_features, bias=None, residual=None, ln_weight=None,
                   ln_bias=None, ln_eps: float = 1e-5, relu: bool = False, out=None,
                   save_for_backward: bool = False, gate=None, row_mask=None, bf16: bool = False,
                   out_dtype=None):
    """y = epilogue(x @ W^T + bias) on the tensor cores (3xTF32); x (rows, in) fp32 contiguous.
    ``bf16``: operands rounded to bf16 on their way into the tensor cores (PCT_LINEAR_BF16);
    ``w_packed`` must then come from ``linear_pack_weight(w, bf16=True)``.  In that mode ``x``, the
    residual / gate and (``out_dtype=torch.bfloat16``) the result may be bf16 tensors
    (PCT_LINEAR_X_BF16 / RES_BF16 / Y_BF16); LayerNorm's saved input and statistics stay fp32.
    With ``save_for_backward`` (LayerNorm epilogue only) returns (y, pre_norm, mean, rstd).
    ``gate`` (rows, out): y is zeroed where gate <= 0 (ReLU backward fused into a dgrad GEMM).
    ``row_mask`` (rows,) bool / uint8: rows with a non-zero entry are written as zeros."""
    rows, k = x.shape
    n = out_features
    flags = (LINEAR_RELU if relu else 0) | (LINEAR_RESIDUAL if residual is not None else 0) | \
        (LINEAR_LAYERNORM if ln_weight is not None else 0) | (LINEAR_GATE if gate is not None else 0) | \
        (LINEAR_BF16 if bf16 else 0)
    if gate is not None:
        if residual is not None:
            raise ValueError("gate and residual share an argument slot of the C ABI")
        residual = gate
    if row_mask is not None:
        if residual is not None or ln_weight is not None:
            raise ValueError("row_mask shares an argument slot of the C ABI with residual / gate")
        if row_mask.numel() != rows or row_mask.element_size() != 1 or not row_mask.is_contiguous():
            raise ValueError("row_mask must be a contiguous (rows,) bool / uint8 tensor")
        residual = row_mask
        flags |= LINEAR_ROWMASK
    if out is not None:
        out_dtype = out.dtype
    elif out_dtype is None:
        out_dtype = torch.float32
    low = [t for t in (x, residual if row_mask is None else None) if t is not None and t.dtype == torch.bfloat16]
    if (low or out_dtype == torch.bfloat16) and not bf16:
        raise TypeError("bf16 tensors need the bf16 operand kernel (bf16=True)")
    if x.dtype == torch.bfloat16:
        flags |= LINEAR_X_BF16
    if row_mask is None and residual is not None and residual.dtype == torch.bfloat16:
        flags |= LINEAR_RES_BF16
    if out_dtype == torch.bfloat16:
        flags |= LINEAR_Y_BF16
    y = out if out is not None else torch.empty((rows, n), dtype=out_dtype, device=x.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    pre = mean = rstd = None
    if save_for_backward:
        pre = torch.empty((rows, n), dtype=torch.float32, device=x.device)
        mean = torch.empty((rows, 1), dtype=torch.float32, device=x.device)
        rstd = torch.empty((rows, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x):
        check(load().pct_linear_forward_train(
            x.data_ptr(), w_packed.data_ptr(), ptr(bias), ptr(residual), ptr(ln_weight),
            ptr(ln_bias), float(ln_eps), y.data_ptr(), ptr(pre), ptr(mean), ptr(rstd), rows, k, n,
            flags, _stream_ptr(x.device)), "pct_linear_forward")
    return (y, pre, mean, rstd) if save_for_backward else y


def layernorm_backward_supported(features: int) -> bool:
    return features % 128 == 0 and 128 <= features <= 512


def layernorm_backward(grad_y, pre_norm, mean, rstd, gamma):
    """One-pass LayerNorm backward: returns (grad_pre (rows, n), grad_gamma (n,), grad_beta (n,))
    for y = (pre_norm - mean) * rstd * gamma + beta over the last dim; fp32 contiguous tensors."""
    rows, n = grad_y.shape
    g = torch.empty_like(grad_y)
    gg = torch.empty((n,), dtype=grad_y.dtype, device=grad_y.device)
    gb = torch.empty((n,), dtype=grad_y.dtype, device=grad_y.device)
    with torch.cuda.device_of(grad_y):
        check(load().pct_linear_layernorm_backward(
            grad_y.data_ptr(), pre_norm.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
            gamma.data_ptr(), g.data_ptr(), gg.data_ptr(), gb.data_ptr(), rows, n,
            _stream_ptr(grad_y.device)), "pct_linear_layernorm_backward")
    return g, gg, gb


def linear_wgrad(grad_y, x, with_bias: bool = False, bf16: bool = False):
    """grad_w (out, in) = grad_y (rows, out)^T @ x (rows, in) on the tensor cores (3xTF32, or with
    ``bf16`` both operands rounded to bf16); ``with_bias`` also returns grad_b (out,) =
    grad_y.sum(0) from the same pass."""
    rows, n = grad_y.shape
    k = x.shape[1]
    gw = torch.empty((n, k), dtype=torch.float32, device=x.device)
    gb = torch.empty((n,), dtype=torch.float32, device=x.device) if with_bias else None
    flags = LINEAR_BF16 if bf16 else 0
    if x.dtype == torch.bfloat16:
        flags |= LINEAR_X_BF16
    if grad_y.dtype == torch.bfloat16:
        flags |= LINEAR_GY_BF16
    if flags & (LINEAR_X_BF16 | LINEAR_GY_BF16) and not bf16:
        raise TypeError("bf16 tensors need the bf16 operand kernel (bf16=True)")
    with torch.cuda.device_of(x):
        check(load().pct_linear_wgrad_ex(grad_y.data_ptr(), x.data_ptr(), gw.data_ptr(),
                                         gb.data_ptr() if with_bias else None, rows, k, n,
                                         flags, _stream_ptr(x.device)),
              "pct_linear_wgrad")
    return (gw, gb) if with_bias else gw


def linear_forward_pair(x, w_packed, out_a, out_b, bias=None, bf16: bool = False):
    """(y_a, y_b) = split of ``x @ W^T + bias`` with W (out_a + out_b, in) packed as one weight: two
    Linears on the same input as one GEMM, two dense results (pct_linear_forward_split)."""
    rows, k = x.shape
    ya = torch.empty((rows, out_a), dtype=torch.float32, device=x.device)
    yb = torch.empty((rows, out_b), dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x):
        check(load().pct_linear_forward_split(
            x.data_ptr(), None, k, 0, w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
            None, ya.data_ptr(), yb.data_ptr(), out_a, out_b, rows, LINEAR_BF16 if bf16 else 0,
            _stream_ptr(x.device)), "pct_linear_forward_split")
    return ya, yb


def linear_forward_cat(x_a, x_b, w_packed, out_features, residual=None, bf16: bool = False):
    """``[x_a | x_b] @ W^T (+ residual)`` without materialising the concatenation: the input gradient of
    two Linears on the same input, W the packed transpose of their row-concatenated weight."""
    rows = x_a.shape[0]
    y = torch.empty((rows, out_features), dtype=torch.float32, device=x_a.device)
    flags = (LINEAR_BF16 if bf16 else 0) | (LINEAR_RESIDUAL if residual is not None else 0)
    with torch.cuda.device_of(x_a):
        check(load().pct_linear_forward_split(
            x_a.data_ptr(), x_b.data_ptr(), x_a.shape[1], x_b.shape[1], w_packed.data_ptr(), None,
            residual.data_ptr() if residual is not None else None, y.data_ptr(), None, out_features, 0,
            rows, flags, _stream_ptr(x_a.device)), "pct_linear_forward_split")
    return y


def linear_wgrad_pair(grad_y_a, grad_y_b, x, bf16: bool = False):
    """(grad_w, grad_b) of the row-concatenated weight / bias of two Linears on the same input."""
    rows, k = x.shape
    na, nb = grad_y_a.shape[1], grad_y_b.shape[1]
    gw = torch.empty((na + nb, k), dtype=torch.float32, device=x.device)
    gb = torch.empty((na + nb,), dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x):
        check(load().pct_linear_wgrad_split(
            grad_y_a.data_ptr(), grad_y_b.data_ptr(), na, nb, x.data_ptr(), gw.data_ptr(), gb.data_ptr(),
            rows, k, LINEAR_BF16 if bf16 else 0, _stream_ptr(x.device)), "pct_linear_wgrad_split")
    return gw, gb

