#!/usr/bin/env python
"""Launches of the bf16-operand Linear kernel for an ncu capture (LayerNorm and ReLU epilogues).
PCT_LINEAR_PERSISTENT=0/1 in the environment picks the one-tile-per-CTA or the persistent kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pctrans_b200 import _cabi  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
rows = 86016
for k, n, mode in ((256, 256, "ln"), (256, 1024, "relu")):
    x = torch.randn(rows, k, device=dev)
    w = torch.randn(n, k, device=dev) / 16
    b = torch.randn(n, device=dev)
    res = torch.randn(rows, n, device=dev) if mode == "ln" else None
    g = torch.ones(n, device=dev) if mode == "ln" else None
    wp = _cabi.linear_pack_weight(w, bf16=True)
    for _ in range(2):
        y = _cabi.linear_forward(x, wp, n, b, res, g, g, 1e-5, relu=(mode == "relu"), bf16=True)
    torch.cuda.synchronize()
print("ok")
