#!/usr/bin/env python
"""Raw PCIe rates of the box with pinned host buffers (what bounds bench.py's `e2e`): H2D alone,
D2H alone, both directions at once, in whole-buffer and 17 MB pieces."""
import torch

dev = torch.device("cuda", 0)
nbytes = 275_251_200
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, pieces):
    step = nbytes // pieces
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(5):
        for p in range(pieces):
            sl = slice(p * step, (p + 1) * step)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return ms, nbytes / ms / 1e6


for pieces in (1, 16, 64):
    for name, a, b in (("H2D only", True, False), ("D2H only", False, True), ("both", True, True)):
        run(a, b, pieces)
        ms, gbs = run(a, b, pieces)
        print(f"pieces {pieces:3d}  {name:9s} {ms:7.3f} ms per 275 MB  = {gbs:5.1f} GB/s per direction")
