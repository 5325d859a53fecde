#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json from the REAL reference: /root/reference/seqoia.h
compiled as-is into oracle/_ref/libsqoa_ref.so (``make -C oracle ref``).  Run in
the build container (the GPU box has no /root/reference); the JSON files are
committed so that every other machine can pin the restatement and the CUDA
kernels to the reference's bytes.

  kat.json      small known-answer vectors: pixels + stream (hex) for encode,
                stream + pixels (or null = rejected) for decode, incl. the
                SURVEY.md appendix C vectors, REF-op streams and rejections.
  digests.json  SHA-256 of the reference's streams / decoded pixels for the
                deterministic synthetic images of seqoia_b200.synth (cfg1, cfg2,
                icons, ...) at full BASELINE.json sizes.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from seqoia_b200 import synth  # noqa: E402


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    oracle.build(with_reference=True)
    ref = oracle.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsqoa_ref.so not built: /root/reference missing?")
    rng = np.random.default_rng(20261018)
    enc, dec = [], []

    def add_enc(name, px, w, h, ch, cs=0, qoi=0):
        px = np.ascontiguousarray(px, dtype=np.uint8).reshape(-1)
        s = ref.encode(px, w, h, ch, cs, qoi)
        enc.append(dict(name=name, w=w, h=h, channels=ch, colorspace=cs, qoi=qoi, pixels=px.tobytes().hex(),
                        stream=None if s is None else s.hex()))
        return s

    def add_dec(name, stream, channels):
        px, d = ref.decode(stream, channels)
        dec.append(dict(name=name, stream=bytes(stream).hex(), channels=channels,
                        pixels=None if px is None else px.tobytes().hex(),
                        desc=[d.width, d.height, d.channels, d.colorspace, d.qoi_compat]))

    # SURVEY.md appendix C
    four = [10, 20, 30, 255, 10, 20, 30, 255, 11, 21, 31, 255, 200, 100, 50, 128]
    for q in (0, 1):
        add_enc(f"surveyC_4x1_q{q}", four, 4, 1, 4, 0, q)
    for n in (2, 61, 62, 63, 64, 122, 123, 512, 513, 514, 1025, 1200):
        for q in (0, 1):
            add_enc(f"solid_{n}_q{q}", [9, 9, 9, 255] * n, n, 1, 4, 0, q)
    for q in (0, 1):
        add_enc(f"start_px_x3_q{q}", [0, 0, 0, 255] * 3, 3, 1, 4, 0, q)
        add_enc(f"zero_px_q{q}", [0, 0, 0, 0], 1, 1, 4, 0, q)
    for ch in (1, 2, 3, 4, 5, 6):
        st = (1 if ch < 3 else 3) + (1 if ch % 2 == 0 else 0)
        first = [50, 60, 70, 80][:st]
        second = list(first)
        second[0] += 1
        add_enc(f"two_px_ch{ch}", first + second, 2, 1, ch, 1, 0)
        add_enc(f"mono_qoi_reject_ch{ch}", first + second, 2, 1, ch, 0, 1)
    add_enc("too_big", [0] * 4, 20000, 20000, 4)
    add_enc("bad_colorspace", [1, 2, 3, 4], 1, 1, 4, 2, 0)
    # a run of 1199 after one pixel, then a different pixel
    add_enc("run_1199", [9, 9, 9, 255] * 1200 + [1, 2, 3, 4], 1201, 1, 4)
    add_enc("run_1199_qoi", [9, 9, 9, 255] * 1200 + [1, 2, 3, 4], 1201, 1, 4, 0, 1)
    # random small images of every kind, all channel counts, both formats
    for k in range(120):
        ch = int(rng.integers(1, 7))
        st = (1 if ch < 3 else 3) + (1 if ch % 2 == 0 else 0)
        w, h = int(rng.integers(1, 40)), int(rng.integers(1, 12))
        n = w * h
        mode = k % 4
        if mode == 0:
            px = rng.integers(0, 256, (n, st))
        elif mode == 1:
            px = (rng.integers(0, 256, (1, st)) + np.cumsum(rng.integers(-3, 4, (n, st)), axis=0)) % 256
        elif mode == 2:
            pal = rng.integers(0, 256, (4, st))
            px = pal[np.resize(np.repeat(rng.integers(0, 4, n), rng.integers(1, 90, n)), n)]
        else:
            pal = rng.integers(0, 256, (20, st))
            px = pal[rng.integers(0, 20, n)]
        s = add_enc(f"rand{k}_m{mode}", px, w, h, ch, k & 1, int(rng.integers(0, 2)))
        if s is not None and k % 3 == 0:
            for oc in (0, 1, 2, 3, 4):
                add_dec(f"rand{k}_dec_c{oc}", s, oc)

    # decoder-only behaviour: REF (SURVEY A.6 probe A), alpha after run / literal, truncation, rejects
    def sq(w, h, ch, body, cs=0):
        return (b"Sqoa" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, cs, 0x31]) + bytes(body)
                + bytes([0, 0, 0, 0, 0, 0, 0, 1]))

    def qf(w, h, ch, body, cs=0):
        return b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, cs]) + bytes(body) + bytes(
            [0, 0, 0, 0, 0, 0, 0, 1])

    probe_a = [0xfe, 0x0a, 0x14, 0x1e, 0xa1, 0x99, 0x00, 0xfe, 0x01, 0x02, 0x03]
    for oc in (0, 3, 4):
        add_dec(f"ref_probeA_c{oc}", sq(6, 1, 3, probe_a), oc)
        add_dec(f"ref_probeA_rgba_c{oc}", sq(6, 1, 4, probe_a), oc)
    add_dec("ref_before_start", sq(3, 1, 4, [0x5f, 0x5f, 0x5f]), 0)
    add_dec("alpha_after_rgb", sq(2, 1, 4, [0xfe, 1, 2, 3, 0x71, 0xc0, 0x62]), 0)
    add_dec("alpha_after_run", sq(4, 1, 4, [0xa0, 0x88, 0xc1, 0x72, 0xc0]), 0)
    add_dec("alpha_as_op_start", sq(70, 1, 4, [0xa0, 0x88, 0x61, 0x65, 0xa0, 0x88]), 0)
    add_dec("truncated_body", sq(9, 2, 4, [0xff, 5, 6, 7, 8]), 0)
    add_dec("bigrun_overshoot", sq(5, 1, 3, [0xfd]), 0)
    add_dec("qoi_index_unwritten", qf(4, 1, 4, [0x05, 0x00, 0x35, 0xc0]), 0)
    add_dec("qoi_starts_with_run", qf(6, 1, 4, [0xc1, 0x35, 0xfe, 1, 2, 3, 0x35]), 0)
    add_dec("qoi_fd_is_run62", qf(70, 1, 3, [0xfd, 0x6a, 0xfd]), 0)
    add_dec("sqoa_magic_no_start_byte", b"Sqoa" + (2).to_bytes(4, "big") + (1).to_bytes(4, "big") + bytes(
        [4, 0, 0xfe, 1, 2, 3, 0xc0, 0, 0, 0, 0, 0, 0, 0, 1]), 0)
    add_dec("qoif_with_start_byte", b"qoif" + (2).to_bytes(4, "big") + (1).to_bytes(4, "big") + bytes(
        [4, 0, 0x31, 0xfe, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 1]), 0)
    add_dec("bad_magic", b"Xqoa" + sq(2, 1, 4, [0xc0])[4:], 0)
    add_dec("too_short", sq(1, 1, 4, [])[:21], 0)
    add_dec("channels_5_arg", sq(1, 1, 4, [0xc0]), 5)
    add_dec("hdr_channels_0", sq(1, 1, 0, [0xc0]), 0)
    add_dec("hdr_channels_6", sq(2, 1, 6, [0xff, 1, 2, 3, 4, 0xc0]), 0)
    add_dec("mono_stream", sq(3, 1, 1, [0xfe, 0x20, 0xa1, 0xc0]), 3)
    add_dec("monoa_stream", sq(3, 1, 2, [0xff, 0x20, 0x40, 0xa1, 0xc0]), 4)
    add_dec("mono_qoi_index128", qf(3, 1, 1, [0xfe, 0x20, 0x7f, 0x10]), 0)
    for k in range(40):
        n = int(rng.integers(8, 60))
        body = rng.integers(0, 256, n, dtype=np.uint8)
        pick = rng.random(n) < 0.6
        body[pick] = rng.choice([0xfe, 0xff, 0xfd, 0xc3, 0x85, 0x65, 0x70, 0x9a, 0x41, 0x05, 0x22], int(pick.sum()))
        ch = int(rng.choice([3, 4]))
        mk = sq if k % 2 == 0 else qf
        add_dec(f"fuzz{k}", mk(int(rng.integers(1, 30)), int(rng.integers(1, 6)), ch, body.tobytes()), int(k % 5))

    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    with open(os.path.join(gdir, "kat.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/libsqoa_ref.so built from /root/reference/seqoia.h (gcc -O3)",
                       encode=enc, decode=dec), f, indent=0)

    # digests of the full-size synthetic configs
    dig = {}

    def add_digest(name, px, w, h, ch, qoi):
        s = ref.encode(px, w, h, ch, 0, qoi)
        back, _ = ref.decode(s, 0)
        assert np.array_equal(back, px.reshape(-1))
        dig[name] = dict(w=w, h=h, channels=ch, qoi=qoi, pixels_sha256=sha(px.tobytes()), stream_len=len(s),
                         stream_sha256=sha(s))

    c1 = synth.cfg1()
    c2 = synth.cfg2()
    c2a = synth.cfg2(channels=4)
    for q in (0, 1):
        add_digest(f"cfg1_1920x1080_rgba_q{q}", c1, 1920, 1080, 4, q)
        add_digest(f"cfg2_3840x2160_rgb_q{q}", c2, 3840, 2160, 3, q)
        add_digest(f"cfg2_3840x2160_rgba_q{q}", c2a, 3840, 2160, 4, q)
    icons = synth.cfg3(256)
    for q in (0, 1):
        h = hashlib.sha256()
        total = 0
        for i in range(256):
            s = ref.encode(icons[i], 64, 64, 4, 0, q)
            h.update(s)
            total += len(s)
        dig[f"cfg3_icons_0_255_q{q}"] = dict(n=256, w=64, h=64, channels=4, qoi=q,
                                              pixels_sha256=sha(icons.tobytes()), stream_len=total,
                                              stream_sha256=h.hexdigest())
    small4 = synth.cfg4(2000, 1999)
    for q in (0, 1):
        add_digest(f"cfg4_scaled_2000x1999_q{q}", small4, 2000, 1999, 4, q)
    scr = synth.image("screen", 1280, 720, 3, seed=7, cell=(160, 90))
    for q in (0, 1):
        add_digest(f"screen_1280x720_rgb_q{q}", scr, 1280, 720, 3, q)
    with open(os.path.join(gdir, "digests.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/libsqoa_ref.so built from /root/reference/seqoia.h (gcc -O3)",
                       generator="seqoia_b200/csrc/synth.c", digests=dig), f, indent=1)
    print(f"kat.json: {len(enc)} encode, {len(dec)} decode vectors; digests.json: {len(dig)} entries")
    for k, v in dig.items():
        print(f"  {k}: {v['stream_len']} B")


if __name__ == "__main__":
    main()
