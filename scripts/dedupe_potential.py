#!/usr/bin/env python
"""How many grad_value reductions could merging save?  (CPU, no GPU needed.)

Counts, on the benchmark's synthetic inputs, the (pixel, head) rows the backward scatters to:
every in-range corner (what the reference does), after merging neighbouring points of a pair (what
msda_bwd_vec_kernel does), after a full de-duplication inside a pair, and after de-duplicating
over the 4 consecutive queries a warp owns (the upper bound of cross-pair merging).

    python scripts/dedupe_potential.py [R|U]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pctrans_b200 import synth
N=2
value, shapes, lsi, loc, attn = synth.make_inputs(N, synth.CVPPP_512, M=8, D=32, P=4, regime=sys.argv[1] if len(sys.argv)>1 else "R", seed=0)
L=3;P=4
tot=0; adj=0; full=0; cross4=0
rows_all=[]
for l,(H,W) in enumerate(synth.CVPPP_512):
    x = loc[...,l,:,0]*W-0.5; y = loc[...,l,:,1]*H-0.5   # (N,Lq,M,P)
    inside=(x>-1)&(y>-1)&(x<W)&(y<H)
    x0=torch.floor(x).long(); y0=torch.floor(y).long()
    rows=[]
    for dy in (0,1):
        for dx in (0,1):
            xx=x0+dx; yy=y0+dy
            ok=inside&(xx>=0)&(xx<W)&(yy>=0)&(yy<H)
            rid=torch.where(ok, yy*W+xx, torch.full_like(xx,-1))
            rows.append(rid)
    rows=torch.stack(rows,-1)  # (N,Lq,M,P,4)
    rows_all.append(rows)
    tot+= (rows>=0).sum().item()
    # full dedupe within pair within level
    r=rows.reshape(-1,P*4)
    rs,_=torch.sort(r,dim=1)
    uniq=((rs[:,1:]!=rs[:,:-1])&(rs[:,1:]>=0)).sum().item()+(rs[:,0]>=0).sum().item()
    full+=uniq
    # adjacent-only: sample p's rows that also appear in sample p-1 are dropped
    a=(rows[...,0,:]>=0).sum().item()
    for p in range(1,P):
        cur=rows[...,p,:]; prev=rows[...,p-1,:]
        dup=(cur.unsqueeze(-1)==prev.unsqueeze(-2)).any(-1)&(cur>=0)
        a+=((cur>=0)&~dup).sum().item()
    adj+=a
    # cross-pair: 4 consecutive queries (same head), full dedupe of rows (upper bound of sector savings needs vector sums)
    Lq=rows.shape[1]
    q4=(Lq//4)*4
    r4=rows[:,:q4].reshape(N,q4//4,4,8,P*4).permute(0,1,3,2,4).reshape(-1,4*P*4)
    rs,_=torch.sort(r4,dim=1)
    cross4+=((rs[:,1:]!=rs[:,:-1])&(rs[:,1:]>=0)).sum().item()+(rs[:,0]>=0).sum().item()
print("taps nonzero-valid",tot,"adjacent-merge",adj, adj/tot,"full within pair",full, full/tot, "warp-level (4 queries) unique", cross4, cross4/tot)
