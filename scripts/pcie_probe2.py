#!/usr/bin/env python
"""Chunked copy patterns of the host pipeline without any compute (GPU box)."""
import torch, time, json
dev = "cuda:0"
N = 16
sizes = {"value": 5376*256*4, "loc": 5376*8*24*4, "attn": 5376*8*12*4, "gout": 5376*256*4}
h_in = {k: torch.empty(N * v, dtype=torch.uint8).pin_memory() for k, v in sizes.items()}
h_out = {k: torch.empty(N * v, dtype=torch.uint8).pin_memory() for k, v in sizes.items()}
d_in = {k: torch.empty(N * v, dtype=torch.uint8, device=dev) for k, v in sizes.items()}
d_out = {k: torch.empty(N * v, dtype=torch.uint8, device=dev) for k, v in sizes.items()}
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
total = sum(sizes.values()) * N
def run(h2d, d2h, chunk, reps=5, dep=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        evs = []
        for lo in range(0, N, chunk):
            if h2d:
                with torch.cuda.stream(s1):
                    for k, v in sizes.items():
                        d_in[k][lo*v:(lo+chunk)*v].copy_(h_in[k][lo*v:(lo+chunk)*v], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s1); evs.append(ev)
            if d2h:
                with torch.cuda.stream(s2):
                    if dep and h2d: s2.wait_event(evs[-1])
                    for k, v in sizes.items():
                        h_out[k][lo*v:(lo+chunk)*v].copy_(d_out[k][lo*v:(lo+chunk)*v], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"ms": round(dt*1e3, 3), "GBps_each": round(total / dt / 1e9, 1)}
run(True, True, 2, 1)
for chunk in (16, 4, 2, 1):
    print(json.dumps({"chunk": chunk, "h2d": run(True, False, chunk), "d2h": run(False, True, chunk),
                      "both": run(True, True, chunk), "both_dep": run(True, True, chunk, dep=True)}))
