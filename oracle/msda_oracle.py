"""CPU oracle for multi-scale deformable attention.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``pctrans_b200/`` does.

Two independent restatements of the reference's result live here
(OPS = /root/reference/connectomics/model/maskformer_block/pixel_decoder/ops):

* ``forward_c`` / ``backward_c`` -- ctypes bindings of ``msda_oracle.c`` (closed form of
  OPS/src/cuda/ms_deform_im2col_cuda.cuh:38-164,258-303).  Fast, fp64 or literal fp32,
  analytic gradients, deterministic summation order.
* ``reference_core`` -- the reference's own ``ms_deform_attn_core_pytorch`` run from the bytecode
  in ``oracle/_ref/`` (built by ``oracle/build_ref.py`` from ``/root/reference``; git-ignored).
* ``core_grid_sample`` -- the algorithm of OPS/functions/ms_deform_attn_func.py:52-72
  (``ms_deform_attn_core_pytorch``): per-level ``F.grid_sample(bilinear, zeros,
  align_corners=False)`` on ``2*loc-1`` followed by the attention-weighted sum.  This is what
  the reference executes on a CPU (the reference has no native CPU kernel,
  OPS/src/cpu/ms_deform_attn_cpu.cpp:22-45), so it is the leg timed as ``cpu_baseline`` and by
  ``bench.py --impl reference`` when ``oracle/_ref/`` is absent (kind "port"; with it the
  reference's own function is timed, kind "reference").

Parity status: pinned.  ``tests/test_oracle.py`` checks both against golden vectors produced
by the reference's own function (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile msda_oracle.c with the committed Makefile (gcc only, no reference sources)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_REF_PYC = os.path.join(_HERE, "_ref", "ms_deform_attn_func.bytecode")
_ref_core = False   # False = not tried yet, None = unavailable


def reference_core():
    """The reference's OWN ``ms_deform_attn_core_pytorch`` (func.py:52-72), executed from the
    bytecode ``oracle/build_ref.py`` compiled out of ``/root/reference`` -- or None when
    ``oracle/_ref/`` was not built (then callers use ``core_grid_sample``, the in-repo port).
    The file imports the reference's CUDA extension at module level (func.py:20-29); an empty
    stand-in module satisfies that import, the CPU function never touches it."""
    global _ref_core
    if _ref_core is not False:
        return _ref_core
    _ref_core = None
    if os.path.exists(_REF_PYC):
        import importlib.util
        import marshal
        import sys
        import types
        sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
        try:
            with open(_REF_PYC, "rb") as f:
                data = f.read()
            if data[:4] != importlib.util.MAGIC_NUMBER:
                raise ImportError("bytecode of another interpreter version")
            mod = types.ModuleType("pctrans_reference_ms_deform_attn_func")
            exec(marshal.loads(data[16:]), mod.__dict__)   # 16-byte .pyc header, then the code object
            _ref_core = mod.ms_deform_attn_core_pytorch
        except Exception:   # treat as not built
            _ref_core = None
    return _ref_core


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        for suffix in ("f64", "f32"):
            f = getattr(lib, f"msda_oracle_forward_{suffix}")
            f.restype = i32
            f.argtypes = [vp] * 6 + [i32] * 7
            g = getattr(lib, f"msda_oracle_backward_{suffix}")
            g.restype = i32
            g.argtypes = [vp] * 9 + [i32] * 7
        _lib = lib
    return _lib


def _np(a, dtype):
    if hasattr(a, "detach"):  # torch tensor
        a = a.detach().cpu()
        if str(a.dtype) == "torch.bfloat16":
            a = a.float()
        a = a.numpy()
    return np.ascontiguousarray(a, dtype=dtype)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2), loc.shape
    return N, S, M, D, L, Lq, P


def forward_c(value, shapes, lsi, loc, attn, dtype=np.float64):
    """out (N,Lq,M*D) as numpy, computed by the C restatement in ``dtype`` arithmetic."""
    lib = _load()
    value, loc, attn = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    shapes, lsi = _np(shapes, np.int64), _np(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    rc = fn(value.ctypes.data, shapes.ctypes.data, lsi.ctypes.data, loc.ctypes.data,
            attn.ctypes.data, out.ctypes.data, N, S, M, D, L, Lq, P)
    if rc != 0:
        raise RuntimeError(f"msda_oracle_forward failed rc={rc}")
    return out


def backward_c(value, shapes, lsi, loc, attn, grad_out, dtype=np.float64):
    """(grad_value, grad_loc, grad_attn) as numpy from the C restatement."""
    lib = _load()
    value, loc, attn = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    grad_out = _np(grad_out, dtype)
    shapes, lsi = _np(shapes, np.int64), _np(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    assert grad_out.size == N * Lq * M * D
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    rc = fn(value.ctypes.data, shapes.ctypes.data, lsi.ctypes.data, loc.ctypes.data,
            attn.ctypes.data, grad_out.ctypes.data, gv.ctypes.data, gl.ctypes.data,
            ga.ctypes.data, N, S, M, D, L, Lq, P)
    if rc != 0:
        raise RuntimeError(f"msda_oracle_backward failed rc={rc}")
    return gv, gl, ga


def core_grid_sample(value, shapes, loc, attn):
    """Torch restatement of ms_deform_attn_core_pytorch (func.py:52-72); differentiable.

    value (N,S,M,D), shapes (L,2) [(H,W)], loc (N,Lq,M,L,P,2) in [0,1], attn (N,Lq,M,L,P)
    -> (N,Lq,M*D).  Same torch ops as the reference (grid_sample bilinear/zeros/
    align_corners=False on 2*loc-1, weighted sum over L*P), written level by level with an
    accumulator instead of the reference's stack+flatten.
    """
    import torch
    import torch.nn.functional as F

    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in (shapes.tolist() if hasattr(shapes, "tolist") else shapes)]
    grid = loc * 2 - 1                                   # func.py:59
    wts = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)   # func.py:70
    acc = None
    start = 0
    for lvl, (H, W) in enumerate(hw):
        v = value[:, start:start + H * W]                # func.py:58 (split by H*W)
        start += H * W
        v = v.permute(0, 2, 3, 1).reshape(N * M, D, H, W)            # func.py:63
        g = grid[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)   # func.py:65
        s = F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False)
        term = (s * wts[:, :, :, lvl]).sum(-1)           # (N*M, D, Lq)
        acc = term if acc is None else acc + term
    return acc.view(N, M * D, Lq).transpose(1, 2).contiguous()      # func.py:71-72
