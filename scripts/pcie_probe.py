#!/usr/bin/env python
"""What can the host link do?  H2D alone, D2H alone, both at once (GPU box)."""
import torch, time, json
dev = "cuda:0"
nbytes = 256 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return nbytes * reps / dt / 1e9
run(True, True, 2)
print(json.dumps({"h2d_only_GBps": run(True, False), "d2h_only_GBps": run(False, True),
                  "both_each_GBps": run(True, True)}))
