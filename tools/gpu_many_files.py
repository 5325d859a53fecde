#!/usr/bin/env python
"""tools/gpu_many_files.py -- what the many-file entry points buy over a loop of sqoa_write / sqoa_read (SURVEY.md 8f:
the file path).  N icons (cfg3) and a few larger images, files on tmpfs.  Usage: gpu_many_files.py [--icons 4096]"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import seqoia_b200 as sb
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--icons", type=int, default=4096)
a = ap.parse_args()
icons = synth.cfg3(a.icons).reshape(a.icons, -1)
big = [synth.image("photo", 1024, 768, 3, seed=k).reshape(-1) for k in range(32)]
imgs = [icons[k] for k in range(a.icons)] + big
descs = [sb.Desc(64, 64, 4, 0, k & 1) for k in range(a.icons)] + [sb.Desc(1024, 768, 3, 0, k & 1) for k in range(32)]
mpx = (a.icons * 4096 + 32 * 1024 * 768) / 1e6
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
with tempfile.TemporaryDirectory(dir=base) as td:
    names = [os.path.join(td, f"f{k}.bin") for k in range(len(imgs))]
    sb.write(names[0], imgs[0], 64, 64, 4, 0, 0)  # warm-up: context, kernels
    t0 = time.perf_counter()
    for nm, im, d in zip(names, imgs, descs):
        sb.write(nm, im, d.width, d.height, d.channels, 0, d.qoi_compat)
    t_loop_w = time.perf_counter() - t0
    ref = [open(nm, "rb").read() for nm in names]
    sb.write_many(names, imgs, descs)  # warm-up at full size: the pinned stage and the device staging grow once
    t0 = time.perf_counter()
    sizes = sb.write_many(names, imgs, descs)
    t_many_w = time.perf_counter() - t0
    same = all(open(nm, "rb").read() == r for nm, r in zip(names, ref)) and all(s == len(r) for s, r in zip(sizes, ref))
    t0 = time.perf_counter()
    loop = [sb.read(nm, 0)[0] for nm in names]
    t_loop_r = time.perf_counter() - t0
    sb.read_many(names, 0)
    t0 = time.perf_counter()
    many = sb.read_many(names, 0)
    t_many_r = time.perf_counter() - t0
    same_r = all(np.array_equal(x, y[0]) for x, y in zip(loop, many)) and all(np.array_equal(x, im) for x, im in zip(loop, imgs))
print(f"{len(imgs)} files ({a.icons} icons + 32 images of 1024x768), {mpx:.1f} Mpx, files on {base or 'tmp'}")
print(f"write: loop of sqoa_write {t_loop_w * 1e3:.1f} ms ({t_loop_w / len(imgs) * 1e6:.0f} us per file)   sqoa_b200_write_many {t_many_w * 1e3:.1f} ms   x{t_loop_w / t_many_w:.1f}   identical files: {same}")
print(f"read:  loop of sqoa_read  {t_loop_r * 1e3:.1f} ms ({t_loop_r / len(imgs) * 1e6:.0f} us per file)   sqoa_b200_read_many  {t_many_r * 1e3:.1f} ms   x{t_loop_r / t_many_r:.1f}   identical pixels: {same_r}")
