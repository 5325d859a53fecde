#!/usr/bin/env python
"""tools/prof_bench_legs.py -- the launches of bench.py's headline step (a batch of 16 cfg2 images per launch, distinct
buffers, far larger than L2), a few of them, for ncu (see profiles/README.md).  --legs picks which."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import seqoia_b200 as sb
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--legs", default="sqoa_encode,sqoa_decode,qoi_encode,qoi_decode")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--images", type=int, default=16)
a = ap.parse_args()
w, h, ch = 3840, 2160, 3
n = a.images
img = synth.cfg2()
raw = w * h * ch
cap = (sb.max_stream_size(w, h, ch) + 63) // 64 * 64
stride = (raw + 63) // 64 * 64
ctx = sb.Context(0)
s = torch.cuda.current_stream().cuda_stream
d_px = torch.empty(n * stride, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(img.reshape(-1)).cuda()
for i in range(n):
    d_px[i * stride: i * stride + raw] = src
legs = a.legs.split(",")
for q, name in ((0, "sqoa"), (1, "qoi")):
    if not any(l.startswith(name) for l in legs):
        continue
    d_st = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_back = torch.empty(n * stride, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(n, dtype=torch.int32, device="cuda")
    ep = ctx.plan([sb.Item(i * stride, i * cap, w, h, 0, ch, 0, q, 0) for i in range(n)])
    reps = a.reps if f"{name}_encode" in legs else 1
    for _ in range(reps):
        ctx.encode_batch(ep, d_px, d_st, d_len, s)
    torch.cuda.synchronize()
    ln = int(d_len[0].item())
    if f"{name}_decode" in legs:
        dp = ctx.plan([sb.Item(i * cap, i * stride, w, h, ln, ch, 0, q, ch) for i in range(n)], decode_=True)
        for _ in range(a.reps):
            ctx.decode_batch(dp, d_st, d_back, d_status, s)
        torch.cuda.synchronize()
        assert torch.equal(d_back.view(n, stride)[:, :raw], d_px.view(n, stride)[:, :raw])
    print(name, "stream bytes per image", ln)
print("launches", ctx.launches)
