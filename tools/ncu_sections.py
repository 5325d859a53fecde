#!/usr/bin/env python
"""tools/ncu_sections.py <report.ncu-rep> -- instruction / stall-sample shares of qoi_rows_kernel by code section
(sections are found by marker comments in qoi_rows_kernels.cuh; inlined helpers from other files are listed by file)."""
import collections
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seqoia_b200", "csrc", "qoi_rows_kernels.cuh")
text = open(src).read().splitlines()
marks = [("helpers", "#pragma once"), ("warp_chain", "SQ_DEV u32 warp_chain"), ("helpers2", "struct ChainHash"),
         ("rows:classify", "SQ_DEV bool rows_pass"), ("rows:scan", "// composition of the ops of this row"),
         ("rows:value+index", "const u32 rgb = x10_to_rgb8(xf);"), ("rows:table", "// the last op of the row with a given hash"),
         ("rows:pixels", "        // pixels"), ("tile:setup", "SQ_DEV void qoi_rows_tile"),
         ("tile:entry walks", "// ---- op boundaries"), ("tile:entry chain", "const int tile_i = (int)t, first_i"),
         ("tile:op walk", "// ---- my true ops"), ("tile:op list", "// the op list, stream order"),
         ("tile:hash+pos chains", "// hash of the running pixel at the tile end"), ("tile:passes+publish", "    RowsOut o;"),
         ("tile:lookback", "// what the table and the running pixel were at my start"),
         ("tile:concretise+patch", "// the table at my start, in shared memory"), ("tile:tail", "if (tv.last_tile) {")]
bounds = []
for name, m in marks:
    for i, l in enumerate(text):
        if m in l:
            bounds.append((i + 1, name))
            break
    else:
        print("marker not found:", m)
bounds.sort()
def section(ln):
    cur = "?"
    for b, n in bounds:
        if ln >= b:
            cur = n
    return cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:."],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
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
agg = collections.Counter()
smp = collections.Counter()
tot = ts = 0
for s in secs:
    ix = {h: i for i, h in enumerate(s["hdr"])}
    ii, si = ix["Instructions Executed"], ix["# Samples"]
    f = s["file"].split("/")[-1]
    for r in s["rows"]:
        if r[2] != "-":
            continue
        try:
            n, sm = float(r[ii]), float(r[si])
        except ValueError:
            continue
        key = section(int(r[0])) if f == "qoi_rows_kernels.cuh" else f
        agg[key] += n
        smp[key] += sm
        tot += n
        ts += sm
print(f"total warp-instructions {tot:.0f}, samples {ts:.0f}")
for k, v in agg.most_common():
    print(f"{v / tot * 100:5.1f}% inst {smp[k] / ts * 100:5.1f}% smp  {k}")
