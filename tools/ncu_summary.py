#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> -- the handful of ncu metrics the roofline discussion needs, per launch."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for w in WANT:
        if w in hdr:
            print(f"  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    stalls = [(float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    for v, n in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {n} = {v:.2f}")
