#!/usr/bin/env python
"""tools/rows_stats.py -- what the tiles of qoi_rows_kernel look like on the BASELINE content (CPU only: runs the
kernel source in the test emulator, which counts per tile: ops, symbolic pixels (patches), second walks, look-back
steps).  Usage: rows_stats.py [rows]   (a stripe of `rows` scanlines of every image, default 160)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle
from seqoia_b200 import synth
from util import Emu

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 160
P = oracle.best()
emu = Emu()
emu.configure(4, 1)


def stats():
    a = (C.c_ulonglong * 13)()
    emu.lib.emu_rows_stats(a)
    return [int(x) for x in a]


cases = [("cfg2 photo RGB", synth.cfg2().reshape(2160, 3840 * 3)[:rows].reshape(-1).copy(), 3840, rows, 3),
         ("cfg2 photo RGBA", synth.cfg2(channels=4).reshape(2160, 3840 * 4)[:rows].reshape(-1).copy(), 3840, rows, 4),
         ("cfg1 mixed RGBA", synth.cfg1().reshape(1080, 1920 * 4)[300:300 + rows].reshape(-1).copy(), 1920, rows, 4),
         ("icon 512 RGBA", synth.image("icon", 512, 512, 4, seed=1001).reshape(-1), 512, 512, 4),
         ("screen RGB", synth.image("screen", 1280, rows, 3, seed=7).reshape(-1), 1280, rows, 3)]
for name, img, w, h, ch in cases:
    s = P.encode(img, w, h, ch, 0, 1)
    stats()
    before = emu.launch_count()
    got, st = emu.decode(s, w * h, ch, 1, ch)
    want, _ = P.decode(s, ch)
    assert np.array_equal(got, want), name
    t, ops, pat, second, steps, *hist = stats()
    print(f"{name:16s} stream {len(s):8d} B  tiles past the first of an image {t:5d}  ops/tile {ops / max(t, 1):6.0f}  "
          f"symbolic pixels/tile {pat / max(t, 1):6.1f}  second walks {second / max(t, 1) * 100:5.1f} %  "
          f"look-back steps/tile {steps / max(t, 1):4.2f}  launches {emu.launch_count() - before}")
    print("                 tiles by symbolic pixels (<=16, 32, 64, 128, 256, 512, 1024, more):", hist)
