#!/usr/bin/env python
"""Where does one encoder layer spend its time?  torch.profiler kernel table (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from pctrans_b200 import encoder as enc, synth

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
train = (sys.argv[2] if len(sys.argv) > 2 else "train") == "train"
torch.manual_seed(0)
model = enc.MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=1,
                                               dim_feedforward=1024, dropout=0.0, num_feature_levels=3).to(dev)
srcs = [torch.randn(n, 256, h, w, device=dev) for h, w in synth.CVPPP_512]
pe = enc.PositionEmbeddingSine(128, normalize=True)
pos = [pe(s) for s in srcs]
amp = len(sys.argv) > 3 and sys.argv[3] == "bf16"
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        if train:
            model.zero_grad(set_to_none=True)
            model(srcs, pos)[0].float().square().mean().backward()
        else:
            with torch.no_grad():
                model(srcs, pos)
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
# every CUDA kernel of the five steps, by total time (no cuBLAS / cuDNN GEMM may appear in the layer)
kern = [e for e in prof.key_averages() if e.device_type.name == "CUDA"] if hasattr(prof.key_averages()[0], "device_type") else []
print("\nCUDA kernels (5 steps):")
for e in sorted(kern, key=lambda e: -e.device_time_total):
    print(f"{e.device_time_total / 5:9.1f} us/step  x{e.count / 5:4.1f}  {e.key[:120]}")
