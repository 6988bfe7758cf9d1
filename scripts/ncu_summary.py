#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): one block of key metrics + top stall reasons per
captured kernel.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [-o profiles/x.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "smsp__inst_executed_op_shfl.sum", "sm__inst_executed_pipe_lsu.sum",
]


def main():
    rep = sys.argv[1]
    out = sys.stdout
    if "-o" in sys.argv:
        out = open(sys.argv[sys.argv.index("-o") + 1], "w")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100, file=out)
        print(name[:160], file=out)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:78s} {r[i]:>18s} {units[i]}", file=out)
        st = []
        for i, h in enumerate(hdr):
            if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("per_issue_active.ratio") \
                    and "not_issued" not in h and r[i] not in ("", "n/a"):
                st.append((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                           float(r[i].replace(",", ""))))
        st.sort(key=lambda x: -x[1])
        print("  stall (warps per issue-active cycle): " + ", ".join(f"{k}={v:.2f}" for k, v in st[:7]), file=out)


if __name__ == "__main__":
    main()
