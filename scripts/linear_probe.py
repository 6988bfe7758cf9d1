#!/usr/bin/env python
"""Correctness and timing probe for the tensor-core Linear (include/pct_linear.h) on a B200.

Compares pct_linear_forward against an fp64 matmul (and torch's fp32 / TF32 GEMMs for scale) on the
encoder's Linear shapes, then times it against torch.nn.functional.linear.  Run under `timeout`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from pctrans_b200 import _cabi  # noqa: E402


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


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


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    out = []
    cases = [(300, 64, 48, "plain"), (128, 16, 16, "plain"), (1000, 256, 256, "plain"),
             (1000, 256, 192, "plain"), (777, 256, 1024, "relu"), (513, 1024, 256, "res"),
             (1000, 256, 256, "ln"), (1000, 128, 128, "ln"), (86016, 256, 256, "ln"),
             (86016, 256, 1024, "relu"), (86016, 1024, 256, "res"), (86016, 256, 288, "plain")]
    for rows, k, n, mode in cases:
        x = torch.randn(rows, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev)
        res = torch.randn(rows, n, device=dev) if mode in ("res", "ln") else None
        g = torch.randn(n, device=dev) if mode == "ln" else None
        be = torch.randn(n, device=dev) if mode == "ln" else None
        wp = _cabi.linear_pack_weight(w)
        y = _cabi.linear_forward(x, wp, n, b, res, g, be, 1e-5, relu=(mode == "relu"))
        torch.cuda.synchronize()
        ref = x.double() @ w.double().t() + b.double()
        if res is not None:
            ref = ref + res.double()
        if mode == "relu":
            ref = ref.relu()
        if mode == "ln":
            ref = F.layer_norm(ref, (n,), g.double(), be.double(), 1e-5)

        def torch_path():
            t = F.linear(x, w, b)
            if res is not None:
                t = t + res
            if mode == "relu":
                t = t.relu()
            if mode == "ln":
                t = F.layer_norm(t, (n,), g, be, 1e-5)
            return t

        torch.backends.cuda.matmul.allow_tf32 = False
        e32 = rel(torch_path(), ref)
        rec = {"rows": rows, "in": k, "out": n, "mode": mode, "err_tc": rel(y, ref), "err_torch_fp32": e32}
        if rows > 10000:
            rec["us_tc"] = timeit(lambda: _cabi.linear_forward(x, wp, n, b, res, g, be, 1e-5, relu=(mode == "relu"), out=y))
            rec["us_torch_fp32"] = timeit(torch_path)
            torch.backends.cuda.matmul.allow_tf32 = True
            rec["err_torch_tf32"] = rel(torch_path(), ref)
            rec["us_torch_tf32"] = timeit(torch_path)
            torch.backends.cuda.matmul.allow_tf32 = False
            rec["tflops_tc"] = 2.0 * rows * k * n / rec["us_tc"] / 1e6
            if k % 32 == 0:
                # bf16 operand precision (PCT_LINEAR_BF16) against torch's bf16 autocast path on the
                # same fp32 tensors (cast kernels + cuBLAS bf16 GEMM + epilogue kernels), and the
                # HBM floor of the fp32 traffic (x, residual, y) at the measured 6 542 GB/s
                wpb = _cabi.linear_pack_weight(w, bf16=True)
                yb = _cabi.linear_forward(x, wpb, n, b, res, g, be, 1e-5, relu=(mode == "relu"), bf16=True)
                rec["err_bf16"] = rel(yb, ref)
                rec["us_bf16"] = timeit(lambda: _cabi.linear_forward(
                    x, wpb, n, b, res, g, be, 1e-5, relu=(mode == "relu"), out=y, bf16=True))

                def autocast_path():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return torch_path()

                rec["err_torch_autocast"] = rel(autocast_path().float(), ref)
                rec["us_torch_autocast"] = timeit(autocast_path)
                traffic = 4.0 * rows * (k + n + (n if res is not None else 0))
                rec["us_hbm_floor"] = traffic / 6542.4e9 * 1e6
            gy = torch.randn(rows, n, device=dev)
            rec["err_wgrad"] = rel(_cabi.linear_wgrad(gy, x), gy.double().t() @ x.double())
            rec["us_wgrad_tc"] = timeit(lambda: _cabi.linear_wgrad(gy, x))
            rec["us_wgrad_torch_fp32"] = timeit(lambda: gy.t() @ x)
            if k % 32 == 0:
                rec["err_wgrad_bf16"] = rel(_cabi.linear_wgrad(gy, x, bf16=True), gy.double().t() @ x.double())
                rec["us_wgrad_bf16"] = timeit(lambda: _cabi.linear_wgrad(gy, x, bf16=True))
        out.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
