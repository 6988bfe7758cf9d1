#!/usr/bin/env python
"""BASELINE configs 2, 4 and 5: the 6-layer deformable-attention encoder around the op.

    python scripts/bench_encoder.py --mode infer --batch 8                      # config 2
    python scripts/bench_encoder.py --mode train --pyramid bbbc --batch 32      # config 4
    torchrun --nproc-per-node N scripts/bench_encoder.py --mode train --global-batch 64   # config 5

One query = one (image, position) pair; a step processes 6*N*Lq query-layers, reported as
queries/s = N*Lq*layers/time.  `--core gridsample` swaps the op for the PyTorch grid_sample
formulation ON THE GPU (the reference's own fallback path, func.py:52-72) for comparison; it is
selected here in the benchmark script only.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from pctrans_b200 import encoder as enc, synth  # noqa: E402
from pctrans_b200.modules import ms_deform_attn as mod_file  # noqa: E402

PYR = {"cvppp512": synth.CVPPP_512, "cvppp448": synth.CVPPP_448, "bbbc": synth.BBBC_520x696}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="train", choices=["infer", "train"])
    ap.add_argument("--pyramid", default="cvppp512", choices=sorted(PYR))
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--global-batch", type=int, default=0, help="total images, split over ranks")
    ap.add_argument("--d-model", type=int, default=256)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--core", default="cuda", choices=["cuda", "gridsample"])
    ap.add_argument("--fused", type=int, default=1, help="use the fused-prologue kernels (default)")
    ap.add_argument("--tc-linears", type=int, default=1,
                    help="fp32 Linears on the tensor cores (3xTF32, functions/linear_tc.py)")
    ap.add_argument("--cuda-graph", type=int, default=0,
                    help="inference only: capture the forward once and time graph replays")
    ap.add_argument("--tf32", type=int, default=0, help="allow TF32 tensor cores for the Linears")
    ap.add_argument("--autocast", default="none", choices=["none", "bf16"],
                    help="run the encoder under bf16 autocast (the reference disables autocast here "
                         "because its kernels are fp32-only, PD/msdeformattn.py:314-320)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.global_batch // world if args.global_batch else args.batch
    pyramid = PYR[args.pyramid]

    if args.core == "gridsample":
        from _gridsample import grid_sample_core

        class _Core:
            @staticmethod
            def apply(value, shapes, lsi, loc, attn, step):
                return grid_sample_core(value, shapes, loc, attn)
        mod_file.MSDeformAttnFunction = _Core

    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    from pctrans_b200.functions import linear_tc
    linear_tc.ENABLED = bool(args.tc_linears)
    torch.manual_seed(0)   # identical initial weights on every rank, as DDP broadcasts them
    model = enc.MSDeformAttnTransformerEncoderOnly(
        d_model=args.d_model, nhead=8, num_encoder_layers=args.layers, dim_feedforward=1024,
        dropout=0.0, num_feature_levels=len(pyramid)).to(dev)
    for mod in model.modules():
        if isinstance(mod, mod_file.MSDeformAttn):
            mod.fused_prologue = bool(args.fused) and args.core == "cuda"
    g = torch.Generator().manual_seed(100 + rank)     # per-rank data (utils/system.py:83)
    srcs = [torch.randn(n, args.d_model, h, w, generator=g).to(dev) for h, w in pyramid]
    pe = enc.PositionEmbeddingSine(args.d_model // 2, normalize=True)
    pos = [pe(s) for s in srcs]
    lq = sum(h * w for h, w in pyramid)
    params = [p for p in model.parameters()]
    side = torch.cuda.Stream()

    import contextlib
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if args.autocast == "bf16" \
        else contextlib.nullcontext

    def step():
        if args.mode == "infer":
            with torch.no_grad(), amp():
                return model(srcs, pos)[0]
        model.zero_grad(set_to_none=True)
        with amp():
            mem = model(srcs, pos)[0]
        mem.float().square().mean().backward()
        if dist is not None:
            # DDP-style exchange: one flat bucket on a side stream (model/build.py:89)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                flat = torch.cat([p.grad.reshape(-1) for p in params])
                dist.all_reduce(flat)
                flat.div_(world)
            torch.cuda.current_stream().wait_stream(side)
        return mem

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    if args.cuda_graph and args.mode == "infer":
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        step = graph.replay
        for _ in range(3):
            step()
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / args.steps
    if rank == 0:
        print(json.dumps({
            "bench": "encoder", "mode": args.mode, "core": args.core, "fused": bool(args.fused),
            "tf32": bool(args.tf32), "tc_linears": bool(args.tc_linears), "cuda_graph": bool(args.cuda_graph), "autocast": args.autocast, "pyramid": args.pyramid,
            "d_model": args.d_model, "layers": args.layers, "n_gpus": world, "batch_per_gpu": n,
            "global_batch": n * world, "ms_per_step": ms,
            "query_layers_per_s": world * n * lq * args.layers / (ms * 1e-3),
            "images_per_s": world * n / (ms * 1e-3),
            "encoder_params": sum(p.numel() for p in params)}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
