#!/usr/bin/env python
"""tools/gpu_shard_passes.py -- where the time of the stream-sharded SQOA decode goes: the three passes (ENTRY, SCAN,
PIXELS) of ONE byte range, timed one by one with CUDA events on one GPU, against the plain decode of the same bytes.
Usage: gpu_shard_passes.py [--rows 5000] [--frac 0.5]   (cfg4 rows; the range is the second `frac` of the stream)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import seqoia_b200 as sb
from seqoia_b200 import dist as sdist
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=5000)
ap.add_argument("--world", type=int, default=2)
a = ap.parse_args()
w, h = 20000, a.rows
img = synth.cfg4_rows(0, h, w, h).reshape(-1)
ctx = sb.Context(0)
s = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(img).cuda()
cap = sb.max_stream_size(w, h, 4)
d_s = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
desc = sb.Desc(w, h, 4, 0, 0)
ctx.encode_device(d_px, desc, d_s, cap, d_n, s)
torch.cuda.synchronize()
n = int(d_n[0].item())
body = n - 23
print(f"{w}x{h} RGBA, stream {n} bytes, {body // 1920} tiles")
d_o = torch.zeros(w * h * 4 + 64, dtype=torch.uint8, device="cuda")
d_st = torch.zeros(4, dtype=torch.int32, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    return ts[len(ts) // 2]


rc, dd, nb = sb.probe(bytes(d_s[:15].cpu().numpy()), n, 0)
t_plain = timed(lambda: ctx.decode_device(d_s, n, dd, 0, d_o, w * h * 4, d_st, s))
print(f"plain decode of the whole stream: {t_plain:.3f} ms")
world = a.world
cuts = sdist.stream_cuts(body, world)
for r in range(world):
    b0, b1 = cuts[r], cuts[r + 1]
    d_body = d_s[15 + b0:]
    avail = min(n - (15 + b0), b1 - b0 + 32)
    d_sum = torch.zeros(8, dtype=torch.int32, device="cuda")
    out = []
    for mode, name in ((sb.DEC_ENTRY, "entry"), (sb.DEC_SCAN, "scan"), (sb.DEC_PIXELS, "pixels")):
        carry = sb.DecCarry(mode, 1 if r else 0, 0, 0, 0xff000000, 1 if r == world - 1 else 0, b1 - b0, 0)
        t = timed(lambda: ctx.decode_shard(d_body, avail, desc, 0, carry, d_sum, d_o, w * h * 4, d_st, s))
        out.append(f"{name} {t:.3f}")
    print(f"range {r} ({(b1 - b0) // 1920} tiles): " + "  ".join(out) + f" ms   (plain / {world} = {t_plain / world:.3f})")
