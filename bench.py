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
`cpu_baseline` = the oracle's port of the reference's CPU path (ms_deform_attn_core_pytorch,
grid_sample) timed on this host.  `--impl reference` times that CPU path alone.
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

    def step():
        value.grad = loc.grad = attn.grad = None
        out = msda_oracle.core_grid_sample(value, shapes, loc, attn)
        out.backward(grad_out)
        return out

    return step, n_images * value.shape[1]


def time_cpu_baseline(pyramid, D, regime, n_images=2, budget_s=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, queries = cpu_step_fn(pyramid, D, n_images, regime)
    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 50):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": queries / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_images} images x {queries // n_images} queries fwd+bwd (torch grid_sample "
                      f"port of ms_deform_attn_core_pytorch, fp32), median of {len(times)} runs, "
                      f"{med * 1e3:.1f} ms/run"}


def run_reference_arm(args, pyramid, pyr_name, D):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs and prints; the others exit 0 without work
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_images = max(1, args.ref_images)
    step, queries = cpu_step_fn(pyramid, D, n_images, args.regime)
    t0 = time.perf_counter()
    step()
    t_one = time.perf_counter() - t0
    if t_one * (args.steps + args.warmup) > 240 and n_images > 1:   # keep the run to minutes
        n_images = 1
        step, queries = cpu_step_fn(pyramid, D, n_images, args.regime)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = queries * args.steps / dt
    sample = (f"{n_images} images x {queries // n_images} queries per step, fwd+autograd bwd of the "
              f"torch grid_sample port of ms_deform_attn_core_pytorch, fp32, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(make_config(args, pyramid, pyr_name, max(1, args.gpus)),
                       cpu_sample_images_per_step=n_images),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
            dist.all_reduce(param_grads)   # the DDP exchange of this path's parameter gradients
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
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
    roofline = {"bound": "hbm", "kernel": f"msda_{'bwd' if dominant == 'backward' else 'fwd'}_vec_kernel",
                "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": ncu_traffic(pyr_name, N, args.dtype, dominant) if args.regime == "R" else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_query": b_bwd if dominant == "backward" else b_fwd,
                "kernel_ms": bwd_ms if dominant == "backward" else fwd_ms,
                "note": "CUDA events around the C-ABI call on the launching stream (backward "
                        "includes the grad_value memset" + (" and the bf16 rounding pass)"
                                                           if args.dtype == "bf16" else ")")}
    step_gbs = queries_per_step * (b_fwd + b_bwd) / (ms_total / args.steps * 1e-3) / 1e9

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
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
