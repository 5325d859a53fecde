#!/usr/bin/env python
"""tools/gpu_decode_probe.py -- bisects a decode problem on the GPU box: every case runs in its own
process under a timeout, so a hang costs seconds and names the failing size."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import oracle, seqoia_b200 as sb
from seqoia_b200 import synth
kind, w, h, ch, qoi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
img = synth.image(kind, w, h, ch, seed=5)
cpu = oracle.best()
s = cpu.encode(img, w, h, ch, 0, qoi)
px, d = sb.decode(s, 0)
ok = px is not None and np.array_equal(px, img.reshape(-1))
print("OK" if ok else "MISMATCH", kind, w, h, ch, qoi, len(s))
''' % ROOT
cases = []
for kind in ("photo", "mixed", "icon"):
    for (w, h) in ((64, 64), (300, 200), (640, 480), (1000, 700), (1920, 1080), (2500, 1601), (3840, 2160)):
        cases.append((kind, w, h, 4, 0))
cases += [("photo", 3840, 2160, 3, 0), ("photo", 640, 480, 3, 1)]
for c in cases:
    try:
        r = subprocess.run([sys.executable, "-c", CASE] + [str(x) for x in c], capture_output=True, text=True, timeout=40)
        print(r.stdout.strip() or ("FAIL " + r.stderr.strip()[-300:]), flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT", c, flush=True)
