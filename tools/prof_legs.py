#!/usr/bin/env python
"""tools/prof_legs.py -- a short, profiler-friendly run: a few launches of each hot kernel on the
cfg2 image (3840x2160 RGB) or cfg1-like RGBA, device resident.  Used under ncu (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import seqoia_b200 as sb
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--legs", default="sqoa_encode,qoi_encode,sqoa_decode,qoi_decode")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--channels", type=int, default=3)
a = ap.parse_args()
w, h, ch = 3840, 2160, a.channels
img = synth.cfg2(channels=ch)
ctx = sb.Context(0)
cap = sb.max_stream_size(w, h, ch)
s = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(img.reshape(-1)).cuda()
legs = a.legs.split(",")
for q, name in ((0, "sqoa"), (1, "qoi")):
    d_s = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
    d_o = torch.zeros(w * h * ch + 64, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(4, dtype=torch.int32, device="cuda")
    desc = sb.Desc(w, h, ch, 0, q)
    reps = a.reps if f"{name}_encode" in legs else 1
    for _ in range(reps):
        ctx.encode_device(d_px, desc, d_s, cap, d_n, s)
    torch.cuda.synchronize()
    n = int(d_n[0].item())
    if f"{name}_decode" in legs:
        rc, dd, nb = sb.probe(bytes(d_s[:15].cpu().numpy()), n, 0)
        for _ in range(a.reps):
            ctx.decode_device(d_s, n, dd, 0, d_o, w * h * ch, d_st, s)
        torch.cuda.synchronize()
        assert torch.equal(d_o[: w * h * ch], d_px)
    print(name, "stream bytes", n)
print("launches", ctx.launches)
