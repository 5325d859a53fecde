#!/usr/bin/env python
"""tools/icon_probe.py -- does the QOI decode pipeline settle on index-heavy icons?  Times the tiled pipeline
(PATH_PARALLEL) against the warp-per-stream kernel on a batch of synthetic icons of one size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seqoia_b200 as sb
from seqoia_b200 import synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
imgs = synth.batch("icon", n, side, side, 4, seed0=0)
px_bytes = side * side * 4
cap = (sb.max_stream_size(side, side, 4) + 63) // 64 * 64
s = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(imgs.reshape(-1)).cuda()
for path in (sb.PATH_PARALLEL, sb.PATH_AUTO):
    ctx = sb.Context(0)
    ctx.set_path(path)
    d_out = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_back = torch.zeros(n * px_bytes, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(n, dtype=torch.int32, device="cuda")
    ep = ctx.plan([sb.Item(i * px_bytes, i * cap, side, side, 0, 4, 0, 1, 0) for i in range(n)])
    ctx.encode_batch(ep, d_px, d_out, d_len, s)
    torch.cuda.synchronize()
    lens = d_len.cpu().numpy()
    dp = ctx.plan([sb.Item(i * cap, i * px_bytes, side, side, int(lens[i]), 4, 0, 1, 4) for i in range(n)], decode_=True)
    for rep in range(3):
        l0 = ctx.launches
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.decode_batch(dp, d_out, d_back, d_status, s)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"side {side} n {n} path {path}: {dt*1e3:.2f} ms, launches {ctx.launches - l0}, ok {bool(torch.equal(d_back, d_px))}, mean stream {lens.mean():.0f} B")
