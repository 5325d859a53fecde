#!/usr/bin/env python
"""tools/gpu_e2e_mt.py -- sqoa_encode / sqoa_decode (pageable buffers) called from T host threads at once: what the
context pool of the host entry points (SQOA_B200_HOST_CONTEXTS) buys.  Usage: gpu_e2e_mt.py --threads 2 --images 16"""
import argparse
import ctypes as C
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import seqoia_b200 as sb
from seqoia_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=2)
ap.add_argument("--images", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
w, h, ch = 3840, 2160, 3
img = np.ascontiguousarray(synth.cfg2(channels=3).reshape(-1)).copy()
L = sb.lib()
L.sqoa_encode.restype = C.c_void_p
L.sqoa_decode.restype = C.c_void_p
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]
ok = [True]


def work(n):
    for _ in range(n):
        for q in (0, 1):
            d = sb.Desc(w, h, ch, 0, q)
            ln = C.c_int(0)
            sp = L.sqoa_encode(img.ctypes.data, C.byref(d), C.byref(ln))
            d2 = sb.Desc()
            pp = L.sqoa_decode(C.c_void_p(sp), ln.value, C.byref(d2), 0)
            if not sp or not pp:
                ok[0] = False
            libc.free(C.c_void_p(sp))
            libc.free(C.c_void_p(pp))


work(1)
best = None
for _ in range(a.reps):
    ts = [threading.Thread(target=work, args=(a.images // a.threads,)) for _ in range(a.threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    best = dt if best is None or dt < best else best
n = a.images // a.threads * a.threads
print(f"threads {a.threads} contexts {os.environ.get('SQOA_B200_HOST_CONTEXTS', '2')} copy threads {os.environ.get('SQOA_B200_COPY_THREADS', 'default')}: "
      f"{4 * n * w * h / best / 1e9:.2f} Gpx/s  ({best / n * 1e3:.2f} ms per image, 4 legs)  ok {ok[0]}")
