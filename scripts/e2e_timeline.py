#!/usr/bin/env python
"""Per-chunk timeline of the host pipeline (GPU box): when does each H2D / compute / D2H run?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pctrans_b200 import synth, _cabi
dev = torch.device("cuda:0")
N, c = 16, 2
value, sh, lsi, loc, attn = synth.make_inputs(N, synth.CVPPP_512, M=8, D=32, P=4, regime="R", seed=0)
gout = torch.randn(N, value.shape[1], 256)
hv, hl, ha, hg = (t.pin_memory() for t in (value, loc, attn, gout))
ho, hgv, hgl, hga = (torch.empty_like(t).pin_memory() for t in (gout, value, loc, attn))
dsh, dlsi = sh.to(dev), lsi.to(dev)
slots = 4
bufs = [dict(v=torch.empty_like(value[:c], device=dev), l=torch.empty_like(loc[:c], device=dev),
             a=torch.empty_like(attn[:c], device=dev), g=torch.empty_like(gout[:c], device=dev),
             o=torch.empty_like(gout[:c], device=dev), gv=torch.empty_like(value[:c], device=dev),
             gl=torch.empty_like(loc[:c], device=dev), ga=torch.empty_like(attn[:c], device=dev)) for _ in range(slots)]
s_in, s_run, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)
def once(record):
    marks = []
    t0 = E(); t0.record(torch.cuda.current_stream())
    for s in (s_in, s_run, s_out): s.wait_stream(torch.cuda.current_stream())
    frees = {}
    for i, lo in enumerate(range(0, N, c)):
        b = bufs[i % slots]; hi = lo + c
        a0, a1, b0, b1, c0, c1 = E(), E(), E(), E(), E(), E()
        with torch.cuda.stream(s_in):
            if i >= slots: s_in.wait_event(frees[i - slots])
            a0.record(s_in)
            b["v"].copy_(hv[lo:hi], non_blocking=True); b["l"].copy_(hl[lo:hi], non_blocking=True)
            b["a"].copy_(ha[lo:hi], non_blocking=True); b["g"].copy_(hg[lo:hi], non_blocking=True)
            a1.record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(a1); b0.record(s_run)
            _cabi.forward(b["v"], dsh, dlsi, b["l"], b["a"], out=b["o"])
            _cabi.backward(b["v"], dsh, dlsi, b["l"], b["a"], b["g"], out=(b["gv"], b["gl"], b["ga"]))
            b1.record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(b1); c0.record(s_out)
            ho[lo:hi].copy_(b["o"], non_blocking=True); hgv[lo:hi].copy_(b["gv"], non_blocking=True)
            hgl[lo:hi].copy_(b["gl"], non_blocking=True); hga[lo:hi].copy_(b["ga"], non_blocking=True)
            c1.record(s_out)
        frees[i] = c1
        marks.append((a0, a1, b0, b1, c0, c1))
    torch.cuda.synchronize()
    if record:
        for i, m in enumerate(marks):
            print(i, " ".join(f"{t0.elapsed_time(e):7.3f}" for e in m))
for _ in range(3): once(False)
print("chunk  h2d_start h2d_end  run_start run_end  d2h_start d2h_end (ms)")
once(True)
