#!/usr/bin/env python
"""tools/make_profile_md.py <report.ncu-rep> <kernel-regex> <title> -- markdown summary of one ncu --set full
capture (launch metrics, stall reasons, hottest source lines) for profiles/."""
import csv
import subprocess
import sys

rep, rx, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"
print(f"# {title}\n")
print(f"Source: `{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on; cold-cache, serialised: compare shares, not absolutes).\n")
import re
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if not re.search(rx, name):
        continue
    print(f"## `{name}`\n")
    print("| metric | value |")
    print("|---|---|")
    for w in WANT:
        if w in hdr:
            print(f"| {w} | {r[hdr.index(w)]} {units[hdr.index(w)]} |")
    stalls = [(float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    print("\nStall reasons (warps per issue-active cycle): " +
          ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]) + "\n")
    break
src = subprocess.run(["python", __file__.replace("make_profile_md.py", "ncu_lines.py"), rep, rx, "0", "18"],
                     capture_output=True, text=True).stdout
print("Hottest source lines (share of executed warp instructions / of stall samples):\n")
print("```")
print("\n".join(l[:150] for l in src.splitlines()[1:]))
print("```")
