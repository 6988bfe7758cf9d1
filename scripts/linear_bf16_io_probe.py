#!/usr/bin/env python
"""Timing of the bf16-operand Linear kernels with fp32 vs bf16 tensors in global memory
(PCT_LINEAR_X_BF16 / Y_BF16 / RES_BF16 / GY_BF16) on the FFN's shapes, 86 016 rows."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pctrans_b200 import _cabi  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


dev = torch.device("cuda", 0)
torch.manual_seed(0)
rows, c, f = 86016, 256, 1024
x = torch.randn(rows, c, device=dev)
h = torch.randn(rows, f, device=dev).relu()
g = torch.randn(rows, c, device=dev)
gh = torch.randn(rows, f, device=dev)
w1 = torch.randn(f, c, device=dev) / 16
w2 = torch.randn(c, f, device=dev) / 32
b1, b2 = torch.randn(f, device=dev), torch.randn(c, device=dev)
p1, p2 = _cabi.linear_pack_weight(w1, bf16=True), _cabi.linear_pack_weight(w2, bf16=True)
p1t = _cabi.linear_pack_weight(w1.t().contiguous(), bf16=True)
p2t = _cabi.linear_pack_weight(w2.t().contiguous(), bf16=True)
h16, gh16 = h.bfloat16(), gh.bfloat16()
B = torch.bfloat16
rec = {
    "linear1 fwd (256->1024 relu), y fp32": timeit(lambda: _cabi.linear_forward(x, p1, f, b1, relu=True, bf16=True)),
    "linear1 fwd, y bf16": timeit(lambda: _cabi.linear_forward(x, p1, f, b1, relu=True, bf16=True, out_dtype=B)),
    "linear2 fwd (1024->256 + res), x fp32": timeit(lambda: _cabi.linear_forward(h, p2, c, b2, x, bf16=True)),
    "linear2 fwd, x bf16": timeit(lambda: _cabi.linear_forward(h16, p2, c, b2, x, bf16=True)),
    "dgrad2 (256->1024, gate), gate fp32 y fp32": timeit(lambda: _cabi.linear_forward(g, p2t, f, gate=h, bf16=True)),
    "dgrad2, gate bf16 y fp32": timeit(lambda: _cabi.linear_forward(g, p2t, f, gate=h16, bf16=True)),
    "dgrad2, gate fp32 y bf16": timeit(lambda: _cabi.linear_forward(g, p2t, f, gate=h, bf16=True, out_dtype=B)),
    "dgrad2, gate bf16 y bf16": timeit(lambda: _cabi.linear_forward(g, p2t, f, gate=h16, bf16=True, out_dtype=B)),
    "dgrad1 (1024->256 + res), x fp32": timeit(lambda: _cabi.linear_forward(gh, p1t, c, residual=g, bf16=True)),
    "dgrad1, x bf16": timeit(lambda: _cabi.linear_forward(gh16, p1t, c, residual=g, bf16=True)),
    "wgrad2 (g, h), fp32 fp32": timeit(lambda: _cabi.linear_wgrad(g, h, with_bias=True, bf16=True)),
    "wgrad2 (g, h), h bf16": timeit(lambda: _cabi.linear_wgrad(g, h16, with_bias=True, bf16=True)),
    "wgrad1 (gh, x), fp32 fp32": timeit(lambda: _cabi.linear_wgrad(gh, x, with_bias=True, bf16=True)),
    "wgrad1 (gh, x), gh bf16": timeit(lambda: _cabi.linear_wgrad(gh16, x, with_bias=True, bf16=True)),
}
for k_, v in rec.items():
    print(f"{k_:50s} {v:8.1f} us")
print(json.dumps(rec))
