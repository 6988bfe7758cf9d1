#!/usr/bin/env python
"""Op-level cost of the prologue: fused kernels vs torch softmax/location ops + plain kernels
(GPU box).  Same inputs, CUDA events, L2 flushed between timed iterations."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pctrans_b200 import synth
from pctrans_b200.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction

dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shapes, M, D, P = synth.CVPPP_512, 8, 32, 4
L = len(shapes)
value, sh, lsi, loc, _ = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime="R", seed=0)
S = value.shape[1]
ref = synth.encoder_reference_points(shapes)[None, :, None, :].expand(N, S, L, 2).contiguous()
norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
off = ((loc - ref[:, :, None, :, None, :]) * norm[None, None, None, :, None, :]).contiguous()
logits = torch.randn(N, S, M, L * P)
gout = torch.randn(N, S, M * D)
v, sh, lsi, ref, off, logits, gout, norm = (t.to(dev) for t in (value, sh, lsi, ref, off, logits, gout, norm))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def fused():
    vv, oo, ll = v.requires_grad_(True), off.requires_grad_(True), logits.requires_grad_(True)
    vv.grad = oo.grad = ll.grad = None
    MSDeformAttnFusedFunction.apply(vv, sh, lsi, ref, oo, ll).backward(gout)

def unfused():
    vv, oo, ll = v.requires_grad_(True), off.requires_grad_(True), logits.requires_grad_(True)
    vv.grad = oo.grad = ll.grad = None
    attn = torch.softmax(ll, -1).view(N, S, M, L, P)
    lc = ref[:, :, None, :, None, :] + oo / norm[None, None, None, :, None, :]
    MSDeformAttnFunction.apply(vv, sh, lsi, lc, attn, 128).backward(gout)

res = {}
for name, fn in (("fused", fused), ("unfused", unfused)):
    for _ in range(5): fn()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name + "_fwd_bwd_ms"] = statistics.median(ts)
res["queries"] = N * S
res["fused_qps"] = N * S / (res["fused_fwd_bwd_ms"] * 1e-3)
res["unfused_qps"] = N * S / (res["unfused_fwd_bwd_ms"] * 1e-3)
print(json.dumps(res))
