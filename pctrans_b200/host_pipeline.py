"""Host-buffer entry to the MSDeformAttn path: pinned host tensors in, pinned host tensors out.

This is the end-to-end call measured as ``e2e`` by bench.py.  Images are independent in this op
(SURVEY.md 8e), so the batch is cut into chunks of images and three CUDA streams overlap the
H2D copy of chunk i+1, the forward+backward kernels of chunk i and the D2H copy of chunk i-1.
Device staging buffers and pinned result buffers are allocated once per shape and reused.
"""
from __future__ import annotations

import torch

from . import _cabi


class HostPipeline:
    def __init__(self, device, dtype=torch.float32, chunk_images: int | None = None, slots: int = 4):
        self.device = torch.device(device)
        self.dtype = dtype
        self.chunk_images = chunk_images
        self.slots = slots
        self._s_in = torch.cuda.Stream(self.device)
        self._s_run = torch.cuda.Stream(self.device)
        self._s_out = torch.cuda.Stream(self.device)
        self._key = None

    def _prepare(self, value, shapes, lsi, loc, attn):
        key = (tuple(value.shape), tuple(loc.shape), tuple(shapes.flatten().tolist()), value.dtype)
        if key == self._key:
            return
        N, S, M, D = value.shape
        Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
        # Chunk schedule: uniform chunks of N/8 images.  Measured on B200 (scripts/e2e_timeline.py,
        # scripts/pcie_probe2.py): the pipeline is bound by the H2D direction of the duplex host
        # link (~41 GB/s while D2H runs at ~47 GB/s); finer or tapered chunks change nothing.
        step = self.chunk_images or max(1, N // 8)
        sizes = [step] * (N // step) + ([N % step] if N % step else [])
        self._sizes = sizes
        c = max(sizes) if sizes else 1
        self._c = c
        dev, dt = self.device, self.dtype
        f32 = torch.float32
        self._d_shapes = shapes.to(dev)
        self._d_lsi = lsi.to(dev)
        mk = lambda shape, d: torch.empty(shape, dtype=d, device=dev)
        self._bufs = []
        for _ in range(self.slots):
            self._bufs.append(dict(
                value=mk((c, S, M, D), dt), loc=mk((c, Lq, M, L, P, 2), f32),
                attn=mk((c, Lq, M, L, P), f32), gout=mk((c, Lq, M * D), dt),
                out=mk((c, Lq, M * D), dt), gvalue=mk((c, S, M, D), dt),
                gloc=mk((c, Lq, M, L, P, 2), f32), gattn=mk((c, Lq, M, L, P), f32),
                ws=torch.empty((S * M * D * c * 4 if dt == torch.bfloat16 else 0,),
                               dtype=torch.uint8, device=dev),
                free=torch.cuda.Event(), loaded=torch.cuda.Event(), done=torch.cuda.Event()))
        pin = lambda shape, d: torch.empty(shape, dtype=d).pin_memory()
        self._h_out = pin((N, Lq, M * D), dt)
        self._h_gvalue = pin((N, S, M, D), dt)
        self._h_gloc = pin((N, Lq, M, L, P, 2), f32)
        self._h_gattn = pin((N, Lq, M, L, P), f32)
        self._key = key

    def forward_backward(self, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                         grad_output):
        """All tensor arguments are HOST tensors (pinned for asynchronous copies).  Returns pinned
        host tensors ``(output, grad_value, grad_sampling_loc, grad_attn_weight)``, valid on
        return."""
        for t in (value, sampling_loc, attn_weight, grad_output):
            if t.is_cuda or not t.is_contiguous():
                raise RuntimeError("HostPipeline takes contiguous host tensors")
        self._prepare(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
        N = value.shape[0]
        s_in, s_run, s_out = self._s_in, self._s_run, self._s_out
        cur = torch.cuda.current_stream(self.device)
        for s in (s_in, s_run, s_out):
            s.wait_stream(cur)
        lo = 0
        for i, n in enumerate(self._sizes):
            hi = lo + n
            b = self._bufs[i % self.slots]
            with torch.cuda.stream(s_in):
                if i >= self.slots:
                    s_in.wait_event(b["free"])       # D2H of the chunk that used this slot is done
                b["value"][:n].copy_(value[lo:hi], non_blocking=True)
                b["loc"][:n].copy_(sampling_loc[lo:hi], non_blocking=True)
                b["attn"][:n].copy_(attn_weight[lo:hi], non_blocking=True)
                b["gout"][:n].copy_(grad_output[lo:hi], non_blocking=True)
                b["loaded"].record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(b["loaded"])
                _cabi.forward(b["value"][:n], self._d_shapes, self._d_lsi, b["loc"][:n],
                              b["attn"][:n], out=b["out"][:n])
                _cabi.backward(b["value"][:n], self._d_shapes, self._d_lsi, b["loc"][:n],
                               b["attn"][:n], b["gout"][:n],
                               out=(b["gvalue"][:n], b["gloc"][:n], b["gattn"][:n]),
                               workspace=b["ws"])
                b["done"].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(b["done"])
                self._h_out[lo:hi].copy_(b["out"][:n], non_blocking=True)
                self._h_gvalue[lo:hi].copy_(b["gvalue"][:n], non_blocking=True)
                self._h_gloc[lo:hi].copy_(b["gloc"][:n], non_blocking=True)
                self._h_gattn[lo:hi].copy_(b["gattn"][:n], non_blocking=True)
                b["free"].record(s_out)
            lo = hi
        s_out.synchronize()
        return self._h_out, self._h_gvalue, self._h_gloc, self._h_gattn
