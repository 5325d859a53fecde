#!/usr/bin/env python
"""tools/gpu_cfg5_flags.py -- which images of the cfg5 corpus leave the one-launch QOI decoder (tuning aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import seqoia_b200 as sb
from seqoia_b200 import synth

ctx = sb.Context(0)
s = torch.cuda.current_stream().cuda_stream
seen = {}
for kind, w, h, ch, seed in synth.cfg5_shapes(float(sys.argv[1]) if len(sys.argv) > 1 else 0.25):
    img = synth.image(kind, w, h, ch, seed=seed).reshape(-1)
    d_px = torch.from_numpy(img).cuda()
    cap = sb.max_stream_size(w, h, ch)
    d_s = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
    d_o = torch.zeros(w * h * ch + 64, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(4, dtype=torch.int32, device="cuda")
    desc = sb.Desc(w, h, ch, 0, 1)
    ctx.encode_device(d_px, desc, d_s, cap, d_n, s)
    torch.cuda.synchronize()
    n = int(d_n[0].item())
    rc, dd, nb = sb.probe(bytes(d_s[:15].cpu().numpy()), n, 0)
    before = ctx.launches
    ctx.decode_device(d_s, n, dd, 0, d_o, w * h * ch, d_st, s)
    torch.cuda.synchronize()
    extra = ctx.launches - before
    ok = bool(torch.equal(d_o[: w * h * ch], d_px))
    key = (kind, ch)
    a = seen.setdefault(key, [0, 0, 0])
    a[0] += 1
    a[1] += extra > 1
    a[2] += not ok
    if extra > 1 and a[1] <= 3:
        print("left the rows kernel:", kind, w, h, ch, "seed", seed, "stream", n, "launches", extra, "ok", ok)
for k, v in seen.items():
    print(k, "images", v[0], "handed on", v[1], "wrong", v[2])
