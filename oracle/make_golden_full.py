#!/usr/bin/env python
"""oracle/make_golden_full.py -- TEST INFRASTRUCTURE ONLY.

SHA-256 digests of the REAL reference's streams (/root/reference/seqoia.h compiled as-is into
oracle/_ref/libsqoa_ref.so) for the BASELINE.json configs at FULL size, where only scaled-down versions were
pinned before: the 20000x19999 RGBA image of cfg4 (the reference's size cap, seqoia.h:470; stream offsets up to
1.99e9), the whole 100,000-icon batch of cfg3 and the 2,851-image corpus of cfg5 (streams concatenated in image
order), both formats each.  Written to tests/golden/digests_full.json; the GPU tests and bench.py compare the
CUDA path's output with them.  Needs ~8 GB of host memory and a few minutes; run in the build container."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from seqoia_b200 import synth  # noqa: E402


def main():
    oracle.build(with_reference=True)
    ref = oracle.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsqoa_ref.so not built: /root/reference missing?")
    out = {}
    t0 = time.time()
    # cfg4 at full size
    w, h = 20000, 19999
    img = synth.cfg4(w, h)
    px_sha = hashlib.sha256(img.reshape(-1).data).hexdigest()
    for q in (0, 1):
        s = ref.encode(img.reshape(-1), w, h, 4, 0, q)
        out[f"cfg4_{w}x{h}_rgba_q{q}"] = dict(w=w, h=h, channels=4, qoi=q, pixels_sha256=px_sha, stream_len=len(s),
                                               stream_sha256=hashlib.sha256(s).hexdigest())
        back, _ = ref.decode(s, 0)
        assert hashlib.sha256(back.data).hexdigest() == px_sha
        print(f"cfg4 q{q}: {len(s)} B ({time.time() - t0:.0f} s)", flush=True)
        del s, back
    del img
    # cfg3: all 100,000 icons, streams concatenated in image order
    n = 100_000
    block = 5000
    hs = {0: hashlib.sha256(), 1: hashlib.sha256()}
    tot = {0: 0, 1: 0}
    hp = hashlib.sha256()
    for first in range(0, n, block):
        icons = synth.cfg3(block, first=first)
        hp.update(icons.reshape(-1).data)
        for q in (0, 1):
            for i in range(block):
                s = ref.encode(icons[i].reshape(-1), 64, 64, 4, 0, q)
                hs[q].update(s)
                tot[q] += len(s)
    for q in (0, 1):
        out[f"cfg3_icons_0_{n - 1}_q{q}"] = dict(n=n, w=64, h=64, channels=4, qoi=q, pixels_sha256=hp.hexdigest(),
                                                  stream_len=tot[q], stream_sha256=hs[q].hexdigest())
    print(f"cfg3: {tot} B ({time.time() - t0:.0f} s)", flush=True)
    # cfg5: the full corpus and the quarter-scale one bench.py used in round 1
    for scale in (1.0, 0.25):
        shapes = synth.cfg5_shapes(scale)
        hs = {0: hashlib.sha256(), 1: hashlib.sha256()}
        tot = {0: 0, 1: 0}
        npx = 0
        for kind, w, h, c, seed in shapes:
            im = synth.image(kind, w, h, c, seed=seed)
            npx += w * h
            for q in (0, 1):
                s = ref.encode(im.reshape(-1), w, h, c, 0, q)
                hs[q].update(s)
                tot[q] += len(s)
        for q in (0, 1):
            out[f"cfg5_scale{scale}_q{q}"] = dict(n=len(shapes), n_px=npx, qoi=q, stream_len=tot[q],
                                                   stream_sha256=hs[q].hexdigest())
        print(f"cfg5 scale {scale}: {len(shapes)} images, {npx / 1e6:.0f} Mpx, {tot} B ({time.time() - t0:.0f} s)", flush=True)
    path = os.path.join(ROOT, "tests", "golden", "digests_full.json")
    with open(path, "w") as f:
        json.dump(dict(source="oracle/_ref/libsqoa_ref.so built from /root/reference/seqoia.h (gcc -O3)",
                       generator="seqoia_b200/csrc/synth.c; oracle/make_golden_full.py", digests=out), f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
