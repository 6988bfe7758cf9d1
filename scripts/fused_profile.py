#!/usr/bin/env python
"""Kernel table of the fused-prologue op (forward + backward), R regime, batch 16 -- and, with
`--plain`, of the plain Function on the same locations / weights for comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from pctrans_b200 import synth
from pctrans_b200.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction

dev = "cuda:0"
N = 16
shapes, M, D, P = synth.CVPPP_512, 8, 32, 4
L = len(shapes)
value, sh, lsi, loc, _ = synth.make_inputs(N, shapes, M=M, D=D, P=P, regime="R", seed=0)
S = value.shape[1]
ref = synth.encoder_reference_points(shapes)[None, :, None, :].expand(N, S, L, 2).contiguous()
norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
off = ((loc - ref[:, :, None, :, None, :]) * norm[None, None, None, :, None, :]).contiguous()
logits = torch.randn(N, S, M, L * P)
gout = torch.randn(N, S, M * D)
v, sh, lsi, ref, off, logits, gout, norm, loc = (t.to(dev) for t in (value, sh, lsi, ref, off, logits, gout, norm, loc))
attn = torch.softmax(logits, -1).view(N, S, M, L, P).contiguous()


def fused():
    vv, oo, ll = v.requires_grad_(True), off.requires_grad_(True), logits.requires_grad_(True)
    vv.grad = oo.grad = ll.grad = None
    MSDeformAttnFusedFunction.apply(vv, sh, lsi, ref, oo, ll).backward(gout)


def plain():
    vv, lc, aa = v.requires_grad_(True), loc.requires_grad_(True), attn.requires_grad_(True)
    vv.grad = lc.grad = aa.grad = None
    MSDeformAttnFunction.apply(vv, sh, lsi, lc, aa, 128).backward(gout)


fn = plain if "--plain" in sys.argv else fused
for _ in range(5):
    fn()
torch.cuda.synchronize()
if "--once" in sys.argv:
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=90))
