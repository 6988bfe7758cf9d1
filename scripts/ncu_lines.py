#!/usr/bin/env python
"""Per-source-line summary of an ncu capture (authoring container, no GPU needed).

    python scripts/ncu_lines.py <report.ncu-rep> <object-or-cubin> <mangled-substring> [--name <demangled-substring>] [--top 40]

`ncu --page source --csv` lists one row per SASS instruction (in address order) without line
numbers; `nvdisasm -g` lists the same instructions with `//## File "...", line N` markers.  This
script zips the two and prints, per source line: executed warp instructions, stall samples and
shared-memory wavefronts -- the view used to decide what to cut next.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj, kernel_sub):
    """[(line_no or None)] per SASS instruction of the first kernel whose name contains kernel_sub."""
    tmp = tempfile.mkdtemp()
    cubins = []
    if obj.endswith(".cubin"):
        cubins = [obj]
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in sorted(os.listdir(tmp)) if f.endswith(".cubin")]
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", "-c", cb], stdout=subprocess.PIPE, text=True).stdout
        out, cur_line, in_fn = [], None, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if in_fn and out:
                    return out
                in_fn = kernel_sub in m.group(1)
                out, cur_line = [], None
                continue
            if not in_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                out.append(cur_line)
        if in_fn and out:
            return out
    raise SystemExit(f"kernel containing {kernel_sub!r} not found in {obj}")


def main():
    rep, obj, ksub = sys.argv[1:4]
    rsub = sys.argv[sys.argv.index("--name") + 1] if "--name" in sys.argv else ksub
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the report may hold several kernels: take the block whose "Kernel Name" row matches
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and rsub in r[1]:
            start = i + 1
            break
    if start is None:
        raise SystemExit("kernel not in report")
    hdr = rows[start]
    body = []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            body.append(r)
    lines = sass_lines(obj, ksub)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} SASS instructions in the object, {len(body)} in the report", file=sys.stderr)
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: defaultdict(float))
    want = ["Instructions Executed", "# Samples", "L1 Wavefronts Shared", "stall_barrier", "stall_long_sb",
            "stall_short_sb", "stall_wait", "stall_mio", "stall_branch_resolving", "stall_math"]
    for ln, r in zip(lines, body):
        for w in want:
            try:
                agg[ln][w] += float(r[col[w]])
            except (KeyError, ValueError):
                pass
    tot = defaultdict(float)
    for d in agg.values():
        for k, v in d.items():
            tot[k] += v
    print("totals:", {k: int(v) for k, v in tot.items()})
    print(f"{'line':>24} {'inst(M)':>8} {'%':>5} {'samples':>8} {'%':>5} {'smemwf(M)':>9}  barrier long_sb short_sb wait mio branch math")
    for ln, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        name = f"{ln[0][:18]}:{ln[1]}" if ln else "?"
        print(f"{name:>24} {d['Instructions Executed'] / 1e6:8.2f} {100 * d['Instructions Executed'] / max(tot['Instructions Executed'], 1):5.1f}"
              f" {int(d['# Samples']):8d} {100 * d['# Samples'] / max(tot['# Samples'], 1):5.1f} {d['L1 Wavefronts Shared'] / 1e6:9.2f}"
              f"  {int(d['stall_barrier']):6d} {int(d['stall_long_sb']):6d} {int(d['stall_short_sb']):6d} {int(d['stall_wait']):5d}"
              f" {int(d['stall_mio']):4d} {int(d['stall_branch_resolving']):5d} {int(d['stall_math']):4d}")


if __name__ == "__main__":
    main()
