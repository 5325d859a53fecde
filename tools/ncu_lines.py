#!/usr/bin/env python
"""tools/ncu_lines.py <report.ncu-rep> [kernel-regex] [index] -- per-CUDA-line instruction and stall-sample
shares from an ncu report captured with --import-source on (read with --print-source sass,cuda)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into (function, file) sections
secs = []
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        secs.append(cur)
    elif r[0] == "Function Name":
        cur["func"] = r[1]
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
funcs = []
for s in secs:
    if s.get("func") not in funcs:
        funcs.append(s.get("func"))
# sections repeat per launch; pick launch `which` of the first function
by_func = {}
for s in secs:
    by_func.setdefault(s["func"], []).append(s)
print("functions:", funcs)
fn = funcs[0] if len(funcs) == 1 else funcs[min(which, len(funcs) - 1)]
lines = {}
tot_i = tot_s = 0
for s in by_func[fn]:
    ix = {h: i for i, h in enumerate(s["hdr"])}
    ii, si = ix["Instructions Executed"], ix["# Samples"]
    for r in s["rows"]:
        if r[2] != "-":  # SASS rows carry an address; CUDA rows have "-"
            continue
        try:
            n, sm = float(r[ii]), float(r[si])
        except ValueError:
            continue
        key = (s["file"].split("/")[-1], r[0], r[1].strip()[:110])
        a = lines.setdefault(key, [0, 0])
        a[0] += n
        a[1] += sm
        tot_i += n
        tot_s += sm
print(fn, "total warp-instructions", tot_i, "samples", tot_s)
for (f, ln, src), (n, sm) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / tot_i * 100:5.1f}% inst {sm / max(tot_s, 1) * 100:5.1f}% smp  {f}:{ln:>4}  {src}")
