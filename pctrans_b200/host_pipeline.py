"""Host-buffer entry to the MSDeformAttn path: host tensors in, pinned host tensors out.

Thin Python face of the C ABI's host pipeline (``pct_msda_host_*`` in include/pct_msda.h,
pctrans_b200/csrc/msda_host.cu): the batch is cut into chunks of images -- the op is per-image
independent -- and three CUDA streams overlap the H2D copy of chunk i+1, the
forward+backward kernels of chunk i and the D2H copy of chunk i-1.  This is the call bench.py
measures as ``e2e``.  Measured on B200 it is bound by the H2D direction of the duplex host link
(~41 GB/s while D2H runs at ~47 GB/s; scripts/e2e_timeline.py, scripts/pcie_probe2.py).
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi


class HostPipeline:
    def __init__(self, device, dtype=torch.float32, chunk_images: int | None = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device; there is no CPU path")
        self.dtype = dtype
        lib = _cabi.load()
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        lib.pct_msda_host_ctx_create.restype = i32
        lib.pct_msda_host_ctx_create.argtypes = [ctypes.POINTER(vp), i32, i32]
        lib.pct_msda_host_ctx_destroy.restype = None
        lib.pct_msda_host_ctx_destroy.argtypes = [vp]
        lib.pct_msda_host_forward.restype = i32
        lib.pct_msda_host_forward.argtypes = [vp] * 7 + [i32] * 8
        lib.pct_msda_host_forward_backward.restype = i32
        lib.pct_msda_host_forward_backward.argtypes = [vp] * 11 + [i32] * 8
        self._lib = lib
        self._ctx = vp()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _cabi.check(lib.pct_msda_host_ctx_create(ctypes.byref(self._ctx), index, chunk_images or 0),
                    "pct_msda_host_ctx_create")
        self._key = None

    def __del__(self):
        ctx = getattr(self, "_ctx", None)
        if ctx:
            self._lib.pct_msda_host_ctx_destroy(ctx)
            self._ctx = None

    @staticmethod
    def _host(t, name):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"HostPipeline takes contiguous host tensors ({name})")
        return t

    def _results(self, value, loc, attn, backward):
        """Pinned result buffers, allocated once per shape."""
        N, S, M, D = value.shape
        key = (tuple(value.shape), tuple(loc.shape), value.dtype, backward)
        if key != self._key:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._out = pin((N, loc.shape[1], M * D), value.dtype)
            if backward:
                self._grads = (pin(tuple(value.shape), value.dtype), pin(tuple(loc.shape), loc.dtype),
                               pin(tuple(attn.shape), attn.dtype))
            self._key = key

    def _dims(self, value, spatial_shapes, loc):
        N, S, M, D = value.shape
        return (N, S, M, D, spatial_shapes.shape[0], loc.shape[1], loc.shape[4],
                _cabi.DTYPE_CODES[value.dtype])

    def forward(self, value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
        """Inference: returns the pinned host output (N, Lq, M*D), valid on return."""
        for t, n in ((value, "value"), (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
            self._host(t, n)
        sh = spatial_shapes.cpu().contiguous()
        lsi = level_start_index.cpu().contiguous()
        self._results(value, sampling_loc, attn_weight, False)
        rc = self._lib.pct_msda_host_forward(
            self._ctx, value.data_ptr(), sh.data_ptr(), lsi.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), self._out.data_ptr(), *self._dims(value, sh, sampling_loc))
        _cabi.check(rc, "pct_msda_host_forward")
        return self._out

    def forward_backward(self, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                         grad_output):
        """All tensor arguments are HOST tensors (pinned for asynchronous full-duplex copies).
        Returns pinned host tensors ``(output, grad_value, grad_sampling_loc, grad_attn_weight)``,
        valid on return and reused by the next call with the same shapes."""
        for t, n in ((value, "value"), (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight"),
                     (grad_output, "grad_output")):
            self._host(t, n)
        sh = spatial_shapes.cpu().contiguous()
        lsi = level_start_index.cpu().contiguous()
        self._results(value, sampling_loc, attn_weight, True)
        gv, gl, ga = self._grads
        rc = self._lib.pct_msda_host_forward_backward(
            self._ctx, value.data_ptr(), sh.data_ptr(), lsi.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), grad_output.data_ptr(), self._out.data_ptr(), gv.data_ptr(),
            gl.data_ptr(), ga.data_ptr(), *self._dims(value, sh, sampling_loc))
        _cabi.check(rc, "pct_msda_host_forward_backward")
        return self._out, gv, gl, ga
