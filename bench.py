#!/usr/bin/env python
"""bench.py -- MSDeformAttn forward+backward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dtype f32|bf16] [--batch B] [--pyramid cvppp512|cvppp448|bbbc] [--regime R|U]

One "step" = one forward + one backward of MSDeformAttnFunction over one batch of synthetic
input (BASELINE config 3 shape by default: CVPPP 512x512 pyramid 16^2/32^2/64^2, d_model 256,
8 heads, 3 levels, 4 points, batch 16 per GPU, realistic sampling locations).  Work is sharded by
image across ranks (weak scaling: every rank owns `--batch` images; no data-path collective --
the op is per-image independent); for N>1 each step also all-reduces a
parameter-gradient-sized buffer over NCCL, which is the only exchange the reference's DDP
training performs for this path.

Prints ONE JSON line on rank 0.  `value` = queries/s with inputs resident in HBM; `e2e` = the
same through the public API with pinned-host inputs/outputs copied inside the timed region;
`roofline` = algorithmic HBM bytes of the dominant kernel / its CUDA-event time / measured peak;
`cpu_baseline` = the reference's own CPU path (ms_deform_attn_core_pytorch, run from the bytecode
in oracle/_ref when it was built, else the oracle's port of it) timed on this host.
`--impl reference` times that CPU path alone.

`sub_records` carries the other BASELINE.json configs, measured in the same run (N=1): config 3 as
stated (bf16 values), the real PCTrans head dim (d_model 128), config 4 (BBBC pyramid, batch 32:
the op alone and the 6-layer encoder fwd+bwd), config 2 (pixel decoder inference, batch 8, checked
against the reference's grid_sample formulation on the same weights) -- each with its time, its
queries/s, the roofline fraction recomputed from that time and the in-run maximum error against
the CPU oracle on one image.  Every run (any N) also measures config 5: the 6-layer encoder
training step at 32 images per GPU (global batch 64 / 128 / 256 at 2 / 4 / 8 GPUs) with the NCCL
all-reduce of its 4.39 M parameter gradients overlapped on a side stream.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "msda_fwd_bwd_queries_per_sec"
UNIT = "queries/s"
PYRAMIDS = {"cvppp512": "CVPPP_512", "cvppp448": "CVPPP_448", "bbbc": "BBBC_520x696"}
M_HEADS, N_POINTS = 8, 4
PARAM_GRAD_FLOATS = 205_600  # four Linears of one MSDeformAttn at d256/L3/P4
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--d-model", type=int, default=256)
    ap.add_argument("--pyramid", default="cvppp512", choices=sorted(PYRAMIDS))
    ap.add_argument("--regime", default="R", choices=["R", "U"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-images", type=int, default=2, help="images per step of the CPU arm")
    ap.add_argument("--no-sub-records", action="store_true",
                    help="skip the other BASELINE configs (sub_records)")
    ap.add_argument("--cfg5-batch", type=int, default=32, help="images per GPU of the config-5 encoder step")
    return ap.parse_args()


def bytes_per_query(D: int, L: int, P: int, val_bytes: int):
    """Compulsory HBM bytes per query (SURVEY.md 8d / BASELINE.md 3): value and grad_value are
    counted once because Lq == S."""
    b_val = M_HEADS * D * val_bytes
    b_loc = M_HEADS * L * P * 2 * 4
    b_att = M_HEADS * L * P * 4
    fwd = 2 * b_val + b_loc + b_att
    bwd = 3 * b_val + 2 * b_loc + 2 * b_att
    return fwd, bwd


def make_config(args, pyramid, pyr_name, world):
    """`config` shared by both arms (the driver compares the lines of the two arms)."""
    L = len(pyramid)
    D = args.d_model // M_HEADS
    S = sum(h * w for h, w in pyramid)
    val_bytes = 4 if args.dtype == "f32" else 2
    b_fwd, b_bwd = bytes_per_query(D, L, N_POINTS, val_bytes)
    return {
        "workload": f"MSDeformAttnFunction fwd+bwd (BASELINE config 3 shape), {pyr_name} "
                    f"pyramid, batch {args.batch}/GPU, d_model {args.d_model}, {M_HEADS} heads, "
                    f"{L} levels, {N_POINTS} points, {args.regime} locations",
        "pyramid": [list(p) for p in pyramid], "batch_per_gpu": args.batch,
        "global_batch": args.batch * world, "queries_per_step_per_gpu": args.batch * S,
        "head_dim": D, "regime": args.regime,
        "parallelism": f"batch-sharded x{world}" + (
            f", NCCL all-reduce of {PARAM_GRAD_FLOATS} fp32 param grads per step" if world > 1 else ""),
        "l2": "no flush: one step touches "
              f"{(b_fwd + b_bwd) * args.batch * S / 1e6:.0f} MB of distinct tensors (> 126 MB L2)",
        # the CPU arm (cpu_baseline / --impl reference) runs this many images per step of the same
        # per-image workload (= BASELINE config 1 at 2); both arms report queries/s
        "cpu_sample_images_per_step": max(1, args.ref_images),
    }


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def onchip_bounds(pyr_name: str, batch: int, dtype: str):
    """The on-chip resources that actually saturate (from the committed ncu capture of the same
    shape, profiles/*_onchip_bounds.json): the HBM roofline is not what limits this gather/scatter
    op.  Informational; None when no capture matches."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_onchip_bounds.json")), reverse=True):
        try:
            with open(path) as f:
                doc = json.load(f)
            entry = doc.get(f"{pyr_name}/{batch}/{dtype}")
            if entry:
                return dict(entry, source=doc.get("source"))
        except Exception:
            continue
    return None


def ncu_traffic(pyr_name: str, batch: int, dtype: str, which: str):
    """DRAM bytes per launch of the forward / backward kernel from the committed ncu --set full
    capture of the same shape (profiles/*_ncu_dram_traffic.json), or None."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_dram_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                entry = json.load(f).get(f"{pyr_name}/{batch}/{dtype}")
            if entry:
                return float(entry[which]["dram_bytes"])
        except Exception:
            continue
    return None


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs local to GPU `index` BEFORE any pinned host buffer is allocated,
    so that the end-to-end arm's host buffers sit in the memory of the socket the GPU hangs off
    (first touch).  With one process per GPU on a multi-socket box every rank otherwise shares
    whichever node the allocator picks, and the ranks' H2D / D2H copies cross the socket link:
    the round-1 e2e number grew only 1.8x from 1 to 8 GPUs.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        return {"cpus_before": before, "cpus_after": len(after), "first_cpu": after[0] if after else None,
                "numa_node": node}
    except Exception as e:   # no NVML / not permitted in this container: run unpinned
        return {"error": f"{type(e).__name__}: {e}"}


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples = []   # (t, sm_mhz, reasons_bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self, t0: float, t1: float):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
                 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                 0x100: "display_clock_setting"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [n for bit, n in names.items() if mask & bit and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(inside)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's port of the reference's CPU path
# --------------------------------------------------------------------------------------------
def cpu_step_fn(pyramid, D, n_images, regime, seed=0):
    from oracle import msda_oracle
    from pctrans_b200 import synth

    value, shapes, lsi, loc, attn = synth.make_inputs(n_images, pyramid, M=M_HEADS, D=D,
                                                      P=N_POINTS, regime=regime, seed=seed)
    grad_out = torch.randn(n_images, value.shape[1], M_HEADS * D,
                           generator=torch.Generator().manual_seed(seed + 1))
    value.requires_grad_(True)
    loc.requires_grad_(True)
    attn.requires_grad_(True)
    ref_core = msda_oracle.reference_core()     # the reference's own function (oracle/_ref), or None
    shape_list = [(int(h), int(w)) for h, w in shapes.tolist()]

    def step():
        value.grad = loc.grad = attn.grad = None
        if ref_core is not None:
            out = ref_core(value, shape_list, loc, attn)
        else:
            out = msda_oracle.core_grid_sample(value, shapes, loc, attn)
        out.backward(grad_out)
        return out

    kind = "reference" if ref_core is not None else "port"
    what = ("the reference's own ms_deform_attn_core_pytorch (bytecode of OPS/functions/"
            "ms_deform_attn_func.py in oracle/_ref)" if ref_core is not None else
            "torch grid_sample port of ms_deform_attn_core_pytorch (oracle/_ref not built)")
    return step, n_images * value.shape[1], kind, what


def time_cpu_baseline(pyramid, D, regime, n_images=2, budget_s=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, queries, kind, what = cpu_step_fn(pyramid, D, n_images, regime)
    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 50):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": queries / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n_images} images x {queries // n_images} queries fwd+autograd bwd of {what}, "
                      f"fp32, median of {len(times)} runs, {med * 1e3:.1f} ms/run"}


def run_reference_arm(args, pyramid, pyr_name, D):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs and prints; the others exit 0 without work
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_images = max(1, args.ref_images)
    step, queries, kind, what = cpu_step_fn(pyramid, D, n_images, args.regime)
    t0 = time.perf_counter()
    step()
    t_one = time.perf_counter() - t0
    if t_one * (args.steps + args.warmup) > 240 and n_images > 1:   # keep the run to minutes
        n_images = 1
        step, queries, kind, what = cpu_step_fn(pyramid, D, n_images, args.regime)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = queries * args.steps / dt
    sample = (f"{n_images} images x {queries // n_images} queries per step, fwd+autograd bwd of "
              f"{what}, fp32, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(make_config(args, pyramid, pyr_name, max(1, args.gpus)),
                       cpu_sample_images_per_step=n_images),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# sub-records: the other BASELINE.json configs, measured in the same run
# --------------------------------------------------------------------------------------------
ENCODER_LAYERS, ENCODER_FFN = 6, 1024


def _rel_err(a, b):
    import numpy as np
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-30))


def op_record(dev, peak, pyr_name, batch, d_model, dtype_name, regime="R", steps=30, warm=5, check=True):
    """MSDeformAttnFunction fwd+bwd at one shape: CUDA-event times of the forward and the backward,
    queries/s, roofline fractions, and the maximum error against the CPU oracle on image 0."""
    from pctrans_b200 import _cabi, synth
    from pctrans_b200.functions import MSDeformAttnFunction

    pyramid = getattr(synth, PYRAMIDS[pyr_name])
    D, L = d_model // M_HEADS, len(pyramid)
    dtype = torch.float32 if dtype_name == "f32" else torch.bfloat16
    value, shapes, lsi, loc, attn = synth.make_inputs(batch, pyramid, M=M_HEADS, D=D, P=N_POINTS,
                                                      regime=regime, seed=7)
    S = value.shape[1]
    gout = torch.randn(batch, S, M_HEADS * D, generator=torch.Generator().manual_seed(8))
    d_sh, d_lsi = shapes.to(dev), lsi.to(dev)
    d_v = value.to(dev, dtype).requires_grad_(True)
    d_lo = loc.to(dev).requires_grad_(True)
    d_at = attn.to(dev).requires_grad_(True)
    d_g = gout.to(dev, dtype)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []
    for k in range(warm + steps):
        d_v.grad = d_lo.grad = d_at.grad = None
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        out = MSDeformAttnFunction.apply(d_v, d_sh, d_lsi, d_lo, d_at, 128)
        e1.record()
        out.backward(d_g)
        e2.record()
        if k >= warm:
            marks.append((e0, e1, e2))
    torch.cuda.synchronize()
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b, _ in marks)
    bwd_ms = statistics.mean(b.elapsed_time(c) for _, b, c in marks)
    q = batch * S
    b_fwd, b_bwd = bytes_per_query(D, L, N_POINTS, 4 if dtype_name == "f32" else 2)
    rec = {
        "workload": f"MSDeformAttnFunction fwd+bwd, {pyr_name} pyramid, batch {batch}, d_model {d_model}, "
                    f"{dtype_name} values, {regime} locations",
        "queries_per_step": q, "ms_per_step": fwd_ms + bwd_ms, "forward_ms": fwd_ms, "backward_ms": bwd_ms,
        "queries_per_s": q / ((fwd_ms + bwd_ms) * 1e-3),
        "roofline_frac": {"forward": q * b_fwd / (fwd_ms * 1e-3) / 1e9 / peak,
                          "backward": q * b_bwd / (bwd_ms * 1e-3) / 1e9 / peak,
                          "step": q * (b_fwd + b_bwd) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak},
        "algorithmic_bytes_per_query": {"forward": b_fwd, "backward": b_bwd},
        "backward_kernel": {0: "generic", 1: "per-tap vector scatter", 2: "cell-binned"}[
            _cabi.backward_plan(batch, S, M_HEADS, D, L, S, N_POINTS, dtype)],
    }
    if check:
        # image 0 against the fp64 C oracle; locations nudged off the pixel grid (where the gradient
        # w.r.t. the location is undefined) exactly as in tests/test_gpu_parity.py
        from oracle import msda_oracle
        v1 = value[:1].to(dtype).float()
        g1 = gout[:1].to(dtype).float()
        lo1 = synth.nudge_off_grid(loc[:1], pyramid)
        at1 = attn[:1]
        o = _cabi.forward(v1.to(dev, dtype), d_sh, d_lsi, lo1.to(dev), at1.to(dev))
        gv, gl, ga = _cabi.backward(v1.to(dev, dtype), d_sh, d_lsi, lo1.to(dev), at1.to(dev), g1.to(dev, dtype))
        torch.cuda.synchronize()
        r_o = msda_oracle.forward_c(v1, shapes, lsi, lo1, at1)
        r_gv, r_gl, r_ga = msda_oracle.backward_c(v1, shapes, lsi, lo1, at1, g1)
        rec["max_err_vs_oracle"] = {"out": _rel_err(o, r_o), "grad_value": _rel_err(gv, r_gv),
                                    "grad_loc": _rel_err(gl, r_gl), "grad_attn": _rel_err(ga, r_ga)}
        rec["tolerance"] = ({"forward": 1e-5, "backward": 1e-4} if dtype_name == "f32"
                            else {"forward": 2e-2, "backward": 2e-2})
        rec["within_tolerance"] = (rec["max_err_vs_oracle"]["out"] <= rec["tolerance"]["forward"] and
                                   max(rec["max_err_vs_oracle"][k] for k in ("grad_value", "grad_loc", "grad_attn"))
                                   <= rec["tolerance"]["backward"])
    return rec


def _encoder(dev, d_model, n_levels, seed=0):
    from pctrans_b200 import encoder as enc
    torch.manual_seed(seed)   # identical initial weights on every rank, as DDP broadcasts them
    return enc.MSDeformAttnTransformerEncoderOnly(
        d_model=d_model, nhead=M_HEADS, num_encoder_layers=ENCODER_LAYERS,
        dim_feedforward=ENCODER_FFN, dropout=0.0, num_feature_levels=n_levels).to(dev)


def encoder_train_record(dev, pyr_name, batch, d_model=256, steps=8, warm=3, dist=None, world=1, rank=0,
                         bf16_autocast=False):
    """BASELINE configs 4 / 5: the 6-layer deformable-attention encoder (PD/msdeformattn.py:23-161),
    forward + backward of a training step; with `dist` the parameter gradients (4.39 M floats at
    d256) are all-reduced over NCCL on a side stream, overlapped with the tail of the backward and
    waited for where the optimizer would consume them.  `bf16_autocast`: the forward runs under
    torch.autocast(bfloat16) -- the Linears' GEMMs (forward, input and weight gradient) take bf16
    operands on the own tensor-core kernels (PCT_LINEAR_BF16); the op stays fp32."""
    import contextlib
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if bf16_autocast else contextlib.nullcontext
    from pctrans_b200 import encoder as enc, synth
    pyramid = getattr(synth, PYRAMIDS[pyr_name])
    model = _encoder(dev, d_model, len(pyramid))
    g = torch.Generator().manual_seed(100 + rank)     # per-rank data (utils/system.py:83)
    srcs = [torch.randn(batch, d_model, h, w, generator=g).to(dev) for h, w in pyramid]
    pe = enc.PositionEmbeddingSine(d_model // 2, normalize=True)
    pos = [pe(x) for x in srcs]
    params = list(model.parameters())
    n_params = sum(p.numel() for p in params)
    reducer = None
    if dist is not None:
        # DDP-style exchange (model/build.py:89): buckets all-reduced on a side stream as soon as
        # backward has produced them (pctrans_b200/dist.py)
        from pctrans_b200.dist import BucketedGradReducer
        reducer = BucketedGradReducer(params, bucket_bytes=3 << 20)

    def step():
        model.zero_grad(set_to_none=True)
        with amp():
            mem = model(srcs, pos)[0]
        mem.float().square().mean().backward()
        if reducer is not None:
            reducer.finish()        # the optimizer step would start here
        return mem

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    lq = sum(h * w for h, w in pyramid)
    n_buckets = len(reducer.buckets) if reducer is not None else 0
    if reducer is not None:
        reducer.remove()
    del model, srcs, pos
    torch.cuda.empty_cache()
    return {
        "workload": f"{ENCODER_LAYERS}-layer MSDeformAttn encoder training step (fwd+bwd), {pyr_name} pyramid, "
                    f"{batch} images/GPU, d_model {d_model}, "
                    + ("bf16 autocast (own tensor-core Linears with bf16 operands forward / dgrad / wgrad, "
                       "fp32 accumulation and tensors except the FFN's hidden activation, fused-prologue "
                       "op on fp32 values)" if bf16_autocast else
                       "fp32 (tensor-core 3xTF32 Linears, fused-prologue op)")
                    + (f", NCCL all-reduce of {n_params} param grads in {n_buckets} buckets on a side "
                       "stream, overlapped with the backward" if dist is not None else ""),
        "n_gpus": world, "batch_per_gpu": batch, "global_batch": batch * world, "ms_per_step": ms,
        "images_per_s": world * batch / (ms * 1e-3),
        "queries_per_s": world * batch * lq * ENCODER_LAYERS / (ms * 1e-3),
        "queries_are": "query-layers: N * Lq * 6 layers per step", "encoder_params": n_params,
    }


def decoder_infer_record(dev, batch=8, steps=10, warm=3, bf16_operands=False):
    """BASELINE config 2: MSDeformAttnPixelDecoder.forward_features (PD/msdeformattn.py:314-360),
    inference, batch 8, CVPPP 512x512 ResNet features; the outputs are compared in the same run with
    the same decoder (same weights) whose op is the reference's grid_sample formulation
    (func.py:52-72) on the GPU.  `bf16_operands`: the decoder pins its encoder to fp32 as the
    reference does (autocast disabled, PD:314-320); with this flag the timed run sets
    linear_tc.PRECISION = "bf16" (bf16 GEMM operands on the own kernels, fp32 tensors and op) and is
    held to BASELINE's bf16 bar against the fp32 reference formulation."""
    from pctrans_b200 import encoder as enc
    from pctrans_b200.functions import linear_tc
    from pctrans_b200.modules import ms_deform_attn as mod_file
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from _gridsample import grid_sample_core

    torch.manual_seed(0)
    shape = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}
    dec = enc.MSDeformAttnPixelDecoder(shape).to(dev).eval()
    g = torch.Generator().manual_seed(5)
    feats = {k: torch.randn(batch, c, 512 // st, 512 // st, generator=g).to(dev) for k, (c, st) in shape.items()}

    def run():
        with torch.no_grad():
            return dec.forward_features(feats)

    linear_tc.PRECISION = "bf16" if bf16_operands else "auto"
    try:
        for _ in range(warm):
            ours = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ours = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        linear_tc.PRECISION = "auto"

    class _RefCore:   # the reference's CPU/debug formulation as the op, on the GPU
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, step):
            return grid_sample_core(value, shapes, loc, attn)

    saved = mod_file.MSDeformAttnFunction
    fused = []
    try:
        mod_file.MSDeformAttnFunction = _RefCore
        for m in dec.modules():
            if isinstance(m, mod_file.MSDeformAttn):
                fused.append((m, m.fused_prologue))
                m.fused_prologue = False
        for _ in range(2):
            ref = run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            ref = run()
        e1.record()
        torch.cuda.synchronize()
        ref_ms = e0.elapsed_time(e1) / 3
    finally:
        mod_file.MSDeformAttnFunction = saved
        for m, f in fused:
            m.fused_prologue = f
    errs = {"mask_features": _rel_err(ours[0], ref[0].double().cpu().numpy()),
            "encoder_lowest_map": _rel_err(ours[1], ref[1].double().cpu().numpy()),
            "multi_scale_max": max(_rel_err(a, b.double().cpu().numpy()) for a, b in zip(ours[2], ref[2]))}
    lq = sum((512 // st) ** 2 for st in (8, 16, 32))
    del dec, feats
    torch.cuda.empty_cache()
    return {
        "workload": f"MSDeformAttnPixelDecoder.forward_features, inference, batch {batch}, 512x512 input "
                    "(res2..res5 features), d_model 256, 6 encoder layers, "
                    + ("bf16 GEMM operands in the encoder's Linears (own kernels), fp32 tensors"
                       if bf16_operands else "fp32"),
        "ms_per_step": ms, "images_per_s": batch / (ms * 1e-3),
        "queries_per_s": batch * lq * ENCODER_LAYERS / (ms * 1e-3),
        "queries_are": "query-layers: N * Lq * 6 layers per forward",
        "same_decoder_with_grid_sample_op_ms": ref_ms,
        "max_err_vs_reference_formulation": errs,
        # the encoder memory within 1e-5; mask_features passes through the FPN's
        # cuDNN convolutions (TF32 by default, as in the reference) whose operand rounding turns
        # the 5e-7 input difference into ~1e-5 noise, so it is held to 1e-4
        "tolerance": ({"encoder_maps": 2e-2, "mask_features": 2e-2} if bf16_operands
                      else {"encoder_maps": 1e-5, "mask_features": 1e-4}),
        "within_tolerance": (max(errs.values()) <= 2e-2 if bf16_operands else
                             (max(errs["encoder_lowest_map"], errs["multi_scale_max"]) <= 1e-5
                              and errs["mask_features"] <= 1e-4)),
    }


def sub_records(dev, peak, args):
    recs = {}

    def guarded(name, fn):
        try:
            recs[name] = fn()
        except Exception as e:   # a sub-record must never take the headline down with it
            recs[name] = {"error": f"{type(e).__name__}: {e}"}

    guarded("cfg3_bf16_op_b16", lambda: op_record(dev, peak, "cvppp512", 16, 256, "bf16"))
    guarded("d128_f32_op_b16", lambda: op_record(dev, peak, "cvppp512", 16, 128, "f32"))
    guarded("cfg4_bbbc_f32_op_b32", lambda: op_record(dev, peak, "bbbc", 32, 256, "f32"))
    guarded("cfg4_bbbc_bf16_op_b32", lambda: op_record(dev, peak, "bbbc", 32, 256, "bf16"))
    guarded("cfg4_bbbc_encoder6_train_b32", lambda: encoder_train_record(dev, "bbbc", 32))
    guarded("cfg4_bbbc_encoder6_train_b32_bf16_autocast",
            lambda: encoder_train_record(dev, "bbbc", 32, bf16_autocast=True))
    guarded("cfg2_decoder_infer_b8", lambda: decoder_infer_record(dev))
    guarded("cfg2_decoder_infer_b8_bf16_operands", lambda: decoder_infer_record(dev, bf16_operands=True))
    return recs


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from pctrans_b200 import synth

    pyr_name = args.pyramid
    pyramid = getattr(synth, PYRAMIDS[pyr_name])
    D = args.d_model // M_HEADS
    L = len(pyramid)

    if args.impl == "reference":
        run_reference_arm(args, pyramid, pyr_name, D)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MSDeformAttn path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    # one process per GPU: pin each to its GPU's socket.  Not at N=1 -- nothing competes for the host
    # links there, and the CPU baseline that runs later in this process must see every core
    affinity = (bind_to_gpu_numa(local_rank)
                if world > 1 and os.environ.get("PCT_BENCH_NO_AFFINITY") != "1" else None)
    dist = None
    if world > 1:
        # NCCL prints "NCCL version ..." (and any NCCL_DEBUG output) to stdout while the
        # communicator is created; rank 0's stdout must carry exactly one JSON line, so file
        # descriptor 1 points at stderr until the first collective has completed
        import torch.distributed as dist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from pctrans_b200 import _cabi
    from pctrans_b200.functions import MSDeformAttnFunction
    from pctrans_b200.host_pipeline import HostPipeline

    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    val_bytes = 4 if args.dtype == "f32" else 2
    N = args.batch
    # seed = rank, as the reference seeds each process (utils/system.py:83)
    value, shapes, lsi, loc, attn = synth.make_inputs(N, pyramid, M=M_HEADS, D=D, P=N_POINTS,
                                                      regime=args.regime, seed=rank)
    S = value.shape[1]
    grad_out = torch.randn(N, S, M_HEADS * D, generator=torch.Generator().manual_seed(1000 + rank))
    host = {"value": value.to(dtype).pin_memory(), "loc": loc.pin_memory(),
            "attn": attn.pin_memory(), "grad_out": grad_out.to(dtype).pin_memory()}
    d_shapes, d_lsi = shapes.to(dev), lsi.to(dev)
    d_value = host["value"].to(dev).requires_grad_(True)
    d_loc = host["loc"].to(dev).requires_grad_(True)
    d_attn = host["attn"].to(dev).requires_grad_(True)
    d_gout = host["grad_out"].to(dev)
    queries_per_step = N * S
    param_grads = torch.zeros(PARAM_GRAD_FLOATS, device=dev) if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None   # the gradient exchange runs beside the next step

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(e_f0=None, e_f1=None, e_b1=None):
        d_value.grad = d_loc.grad = d_attn.grad = None
        if e_f0 is not None:
            e_f0.record()
        out = MSDeformAttnFunction.apply(d_value, d_shapes, d_lsi, d_loc, d_attn, 128)
        if e_f1 is not None:
            e_f1.record()
        out.backward(d_gout)
        if e_b1 is not None:
            e_b1.record()
        if param_grads is not None:
            # the DDP exchange of this path's parameter gradients (model/build.py:89): issued on a
            # side stream once the backward has produced them, so it overlaps the next step's
            # kernels; the timed region ends only after the last exchange has completed
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dist.all_reduce(param_grads)
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)
    barrier()

    # ---- device-resident timed region ------------------------------------------------------
    marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
    launches0 = _cabi.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    e_start, e_end = ev(), ev()
    e_start.record()
    for k in range(args.steps):
        step(*marks[k])
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)
    e_end.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    ms_total = e_start.elapsed_time(e_end)
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b, _ in marks)
    bwd_ms = statistics.mean(b.elapsed_time(c) for _, b, c in marks)
    clocks = sampler.summary(t_wall0, t_wall1)

    t = torch.tensor([ms_total, fwd_ms, bwd_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks
    ms_total, fwd_ms, bwd_ms = (float(x) for x in t.tolist())
    value_qps = world * queries_per_step * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public host API (pinned host in, pinned host out) ------------
    e2e = None
    if not args.no_e2e:
        pipe = HostPipeline(dev, dtype)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            pipe.forward_backward(host["value"], shapes, lsi, host["loc"], host["attn"], host["grad_out"])
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(e2e_steps):
            res = pipe.forward_backward(host["value"], shapes, lsi, host["loc"], host["attn"],
                                        host["grad_out"])
        e1.record()
        barrier()
        ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        h2d = sum(host[k].numel() * host[k].element_size() for k in host)
        d2h = sum(r.numel() * r.element_size() for r in res)
        e2e = {"value": world * queries_per_step * e2e_steps / (float(ms_e2e) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "api": "pct_msda_host_forward_backward (C ABI, include/pct_msda.h) through "
                      "pctrans_b200.host_pipeline.HostPipeline: pinned host buffers in, chunked "
                      "H2D || fwd+bwd kernels || D2H on three streams, pinned host buffers out"}
    sampler.stop()

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    b_fwd, b_bwd = bytes_per_query(D, L, N_POINTS, val_bytes)
    fwd_gbs = queries_per_step * b_fwd / (fwd_ms * 1e-3) / 1e9
    bwd_gbs = queries_per_step * b_bwd / (bwd_ms * 1e-3) / 1e9
    dominant = "backward" if bwd_ms >= fwd_ms else "forward"
    dom_gbs = bwd_gbs if dominant == "backward" else fwd_gbs
    bwd_kernel = {0: "msda_bwd_generic_kernel", 1: "msda_bwd_vec_kernel", 2: "msda_bwd_binned_kernel"}[
        _cabi.backward_plan(N, S, M_HEADS, D, L, S, N_POINTS, dtype)]
    roofline = {"bound": "hbm", "kernel": bwd_kernel if dominant == "backward" else "msda_fwd_vec_kernel",
                "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": ncu_traffic(pyr_name, N, args.dtype, dominant) if args.regime == "R" else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_query": b_bwd if dominant == "backward" else b_fwd,
                "kernel_ms": bwd_ms if dominant == "backward" else fwd_ms,
                "note": "CUDA events around the C-ABI call on the launching stream (backward "
                        "includes the grad_value memset" + (" and the bf16 rounding pass)"
                                                           if args.dtype == "bf16" else ")")}
    step_gbs = queries_per_step * (b_fwd + b_bwd) / (ms_total / args.steps * 1e-3) / 1e9

    # free the headline's tensors before the other configs run
    del d_value, d_loc, d_attn, d_gout, marks
    torch.cuda.empty_cache()
    subs = {}
    if not args.no_sub_records:
        # config 5 on every rank (it is the scaling record); the single-GPU configs at N=1 only
        try:
            subs["cfg5_encoder6_train"] = encoder_train_record(
                dev, "cvppp512", args.cfg5_batch, dist=dist, world=world, rank=rank)
        except Exception as e:
            subs["cfg5_encoder6_train"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            subs["cfg5_encoder6_train_bf16_autocast"] = encoder_train_record(
                dev, "cvppp512", args.cfg5_batch, dist=dist, world=world, rank=rank, bf16_autocast=True)
        except Exception as e:
            subs["cfg5_encoder6_train_bf16_autocast"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            subs.update(sub_records(dev, peak, args))

    try:
        os.sched_setaffinity(0, all_cpus)
    except OSError:
        pass
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = time_cpu_baseline(pyramid, D, args.regime, n_images=args.ref_images)
        line = {
            "metric": METRIC, "value": value_qps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": make_config(args, pyramid, pyr_name, world),
            "roofline": roofline,
            "forward": {"ms": fwd_ms, "queries_per_s": queries_per_step / (fwd_ms * 1e-3),
                        "algorithmic_gbs": fwd_gbs, "frac": fwd_gbs / peak},
            "backward": {"ms": bwd_ms, "queries_per_s": queries_per_step / (bwd_ms * 1e-3),
                         "algorithmic_gbs": bwd_gbs, "frac": bwd_gbs / peak},
            "step_algorithmic_gbs": step_gbs, "step_frac_of_peak": step_gbs / peak,
            "onchip_bounds": onchip_bounds(pyr_name, N, args.dtype) if args.regime == "R" else None,
            "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "host_affinity": affinity,
            "sub_records": subs,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
