#!/usr/bin/env python
"""tools/time_legs.py -- CUDA-event timing of single legs on device-resident buffers (tuning aid, not the bench).
Inputs rotate over enough replicas to exceed L2.  Usage: time_legs.py [--shape 4k3|4k4|big4|big3] [--legs ...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import seqoia_b200 as sb
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4k3")
ap.add_argument("--legs", default="sqoa_encode,sqoa_decode,qoi_encode,qoi_decode")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
if a.shape == "4k3":
    w, h, ch = 3840, 2160, 3
    img = synth.cfg2(channels=3)
elif a.shape == "4k4":
    w, h, ch = 3840, 2160, 4
    img = synth.cfg2(channels=4)
elif a.shape == "1080p4":
    w, h, ch = 1920, 1080, 4
    img = synth.cfg1()
else:  # big: tile the 4K image 12 times vertically (99.5 Mpx)
    ch = 4 if a.shape.endswith("4") else 3
    w, h = 3840, 2160 * 12
    img = np.tile(synth.cfg2(channels=ch).reshape(2160, 3840 * ch), (12, 1))
npx = w * h
raw = npx * ch
reps = a.reps
n_rep = max(2, int(400e6 // raw) + 1)
ctx = sb.Context(0)
cap = sb.max_stream_size(w, h, ch)
s = torch.cuda.current_stream().cuda_stream
d_px = [torch.from_numpy(np.ascontiguousarray(img).reshape(-1)).cuda() for _ in range(n_rep)]
legs = a.legs.split(",")
print(f"shape {w}x{h}x{ch} ({npx/1e6:.1f} Mpx), {n_rep} replicas, lib {sb.LIB_PATH}")
for q, name in ((0, "sqoa"), (1, "qoi")):
    if not any(l.startswith(name) for l in legs):
        continue
    d_s = [torch.zeros(cap + 64, dtype=torch.uint8, device="cuda") for _ in range(n_rep)]
    d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
    d_o = [torch.zeros(raw + 64, dtype=torch.uint8, device="cuda") for _ in range(n_rep)]
    d_st = torch.zeros(4, dtype=torch.int32, device="cuda")
    desc = sb.Desc(w, h, ch, 0, q)
    for k in range(n_rep):
        ctx.encode_device(d_px[k], desc, d_s[k], cap, d_n, s)
    torch.cuda.synchronize()
    n = int(d_n[0].item())
    alg = raw + n
    if f"{name}_encode" in legs:
        for k in range(3):
            ctx.encode_device(d_px[k % n_rep], desc, d_s[k % n_rep], cap, d_n, s)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for k in range(reps):
            ctx.encode_device(d_px[k % n_rep], desc, d_s[k % n_rep], cap, d_n, s)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
        med = ts[len(ts) // 2]
        print(f"{name}_encode: median {med*1e3:.1f} us  min {ts[0]*1e3:.1f} us  {npx/med/1e3:.0f} Mpx/s  {alg/med/1e6:.0f} GB/s  stream {n}")
    if f"{name}_decode" in legs:
        rc, dd, nb = sb.probe(bytes(d_s[0][:15].cpu().numpy()), n, 0)
        for k in range(3):
            ctx.decode_device(d_s[k % n_rep], n, dd, 0, d_o[k % n_rep], raw, d_st, s)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for k in range(reps):
            ctx.decode_device(d_s[k % n_rep], n, dd, 0, d_o[k % n_rep], raw, d_st, s)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
        med = ts[len(ts) // 2]
        ok = torch.equal(d_o[0][:raw], d_px[0])
        print(f"{name}_decode: median {med*1e3:.1f} us  min {ts[0]*1e3:.1f} us  {npx/med/1e3:.0f} Mpx/s  {alg/med/1e6:.0f} GB/s  roundtrip {ok}")
