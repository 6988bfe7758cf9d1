#!/usr/bin/env python
"""Phase timeline of the Linear forward kernel (variant build with -DPCT_LINEAR_TRACE): per traced
CTA (one in 64) the SM clock at start / set-up done / first K step stored / last K step stored /
accumulator complete / epilogue done.  Prints the median duration of each phase in cycles and the
share of a tile's lifetime.  Build + run:  python scripts/linear_trace.py  (on a B200)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pctrans_b200 import build  # noqa: E402

if "PCTRANS_MSDA_LIB" not in os.environ:
    lib = build.build_variant("trace", ["-DPCT_LINEAR_TRACE=1"], sources=("linear_tc.cu",))
    if "--build-only" in sys.argv:
        print(lib)
        sys.exit(0)
    os.environ["PCTRANS_MSDA_LIB"] = lib
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pctrans_b200 import _cabi  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
rows = 86016
lib = _cabi.load()
lib.pct_linear_debug_trace.restype = ctypes.c_int
lib.pct_linear_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = ["set-up", "first step stored", "main loop (stores)", "last MMAs", "epilogue"]
for k, n, mode in ((256, 256, "ln"), (256, 1024, "relu"), (1024, 256, "res"), (256, 256, "plain")):
    for bf16 in (False, True):
        x = torch.randn(rows, k, device=dev)
        w = torch.randn(n, k, device=dev) / 16
        b = torch.randn(n, device=dev)
        res = torch.randn(rows, n, device=dev) if mode in ("ln", "res") else None
        g = torch.ones(n, device=dev) if mode == "ln" else None
        wp = _cabi.linear_pack_weight(w, bf16=bf16)
        for _ in range(3):
            _cabi.linear_forward(x, wp, n, b, res, g, g, 1e-5, relu=(mode == "relu"), bf16=bf16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y = torch.empty(rows, n, device=dev)
        e0.record()
        for _ in range(20):
            _cabi.linear_forward(x, wp, n, b, res, g, g, 1e-5, relu=(mode == "relu"), bf16=bf16, out=y)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        ctas = ((n + 255) // 256) * ((rows + 127) // 128)
        traced = min(4096, (ctas + 63) // 64)
        buf = np.zeros(8 * traced, dtype=np.int64)
        rc = lib.pct_linear_debug_trace(buf.ctypes.data, buf.size)
        assert rc == 0, rc
        t = buf.reshape(traced, 8)[:, :6]
        d = np.diff(t, axis=1)
        life = t[:, 5] - t[:, 0]
        med = np.median(d, axis=0)
        print(f"{k}->{n} {mode} {'bf16' if bf16 else 'tf32x3'}: {us:.1f} us; tile lifetime median {np.median(life):.0f} cycles; "
              + "; ".join(f"{nm} {m:.0f} ({100 * m / np.median(life):.0f} %)" for nm, m in zip(names, med)))
