#!/usr/bin/env python
"""End-to-end (pinned host in / out) throughput of HostPipeline for several chunk sizes (GPU box)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pctrans_b200 import synth
from pctrans_b200.host_pipeline import HostPipeline
dev = torch.device("cuda:0")
N = 16
value, sh, lsi, loc, attn = synth.make_inputs(N, synth.CVPPP_512, M=8, D=32, P=4, regime="R", seed=0)
gout = torch.randn(N, value.shape[1], 256)
host = [t.pin_memory() for t in (value, loc, attn, gout)]
for chunk in (None, 1, 2, 4, 16):
    pipe = HostPipeline(dev, torch.float32, chunk_images=chunk)
    for _ in range(3):
        pipe.forward_backward(host[0], sh, lsi, host[1], host[2], host[3])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        pipe.forward_backward(host[0], sh, lsi, host[1], host[2], host[3])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"chunk_images": chunk, "ms": dt * 1e3,
                      "Mq_per_s": N * value.shape[1] / dt / 1e6}))
