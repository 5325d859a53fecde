#!/usr/bin/env python
"""tools/gpu_e2e.py -- sqoa_encode / sqoa_decode on host buffers, per-call wall times (tuning aid for the host path).
Usage: gpu_e2e.py [--pinned] [--reps N] [--shape 4k3|1080p4|big3]"""
import argparse, ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import seqoia_b200 as sb
from seqoia_b200 import synth
ap = argparse.ArgumentParser()
ap.add_argument("--pinned", action="store_true")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--shape", default="4k3")
a = ap.parse_args()
if a.shape == "4k3": w, h, ch, img = 3840, 2160, 3, synth.cfg2()
elif a.shape == "1080p4": w, h, ch, img = 1920, 1080, 4, synth.cfg1()
else:
    ch = 3; w, h = 3840, 2160 * 12
    img = np.tile(synth.cfg2().reshape(2160, 3840 * 3), (12, 1))
L = sb.lib()
px = np.ascontiguousarray(img).reshape(-1).copy()
if a.pinned:
    import torch
    t = torch.from_numpy(px).pin_memory(); ptr = t.data_ptr()
else:
    ptr = px.ctypes.data
npx = w * h
for q in (0, 1):
    te, td = [], []
    for i in range(a.reps):
        d = sb.Desc(w, h, ch, 0, q); n = C.c_int(0)
        t0 = time.perf_counter(); sp = L.sqoa_encode(ptr, C.byref(d), C.byref(n)); t1 = time.perf_counter()
        d2 = sb.Desc(); pp = L.sqoa_decode(sp, n.value, C.byref(d2), 0); t2 = time.perf_counter()
        ok = np.array_equal(np.frombuffer(C.string_at(pp, npx * ch), dtype=np.uint8), px) if i == 0 else True
        t3 = time.perf_counter(); L._free(sp); L._free(pp); t4 = time.perf_counter()
        te.append(t1 - t0); td.append(t2 - t1)
        if i == 0: print("q", q, "len", n.value, "roundtrip", ok, "free us", round((t4 - t3) * 1e6))
    te, td = sorted(te[1:]), sorted(td[1:])
    print(f"q{q} encode median {te[len(te)//2]*1e6:.0f} us  decode median {td[len(td)//2]*1e6:.0f} us  -> {npx/te[len(te)//2]/1e6:.0f} / {npx/td[len(td)//2]/1e6:.0f} Mpx/s")
