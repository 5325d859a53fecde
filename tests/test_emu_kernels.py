"""Kernel LOGIC on the CPU: the product's kernel source compiled with -DSQ_EMU
against a lock-step warp emulator (tests/emu), compared byte for byte with the
oracle.  Covers what a GPU run cannot be steered into cheaply: several thread
blocks interleaved in different orders (decoupled look-back over aggregate and
inclusive states), tile and row edges, run caps, batches."""
import os

import numpy as np
import pytest

import oracle
from seqoia_b200 import synth
from util import ROOT, Emu, first_difference, golden, random_image, stored_channels


@pytest.fixture(scope="module")
def emu():
    return Emu()


def test_serial_kernels_match_golden_vectors(emu):
    kat = golden("kat.json")
    for v in kat["encode"]:
        if v["stream"] is None:
            continue
        px = np.frombuffer(bytes.fromhex(v["pixels"]), dtype=np.uint8)
        got = emu.serial_encode(px, v["w"], v["h"], v["channels"], v["qoi"], v["colorspace"])
        assert got == bytes.fromhex(v["stream"]), v["name"]
    for v in kat["decode"]:
        if v["pixels"] is None and v["name"] != "ref_before_start":
            continue
        s = bytes.fromhex(v["stream"])
        w, h, hc, _cs, q = v["desc"]
        oc = v["channels"] or stored_channels(hc)
        px, st = emu.serial_decode(s, w, h, hc, q, oc)
        if v["pixels"] is None:
            assert st != 0, v["name"]
        else:
            assert px.tobytes() == bytes.fromhex(v["pixels"]), v["name"]


@pytest.mark.parametrize("qoi", [0, 1])
@pytest.mark.parametrize("ch", [3, 4])
def test_parallel_encoder_matches_oracle(emu, qoi, ch):
    P = oracle.best()
    rng = np.random.default_rng(100 + 2 * ch + qoi)
    for it in range(120):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        if it % 5 == 0:
            w, h = 1024, int(rng.integers(1, 5))
        if it % 9 == 0:
            w, h = 2048 + int(rng.integers(0, 3)), 3
        img = random_image(rng, w * h, ch, it % 4)
        if it % 13 == 0:
            img[:] = img[0]
        emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 12345)
        got = emu.encode(img, w, h, ch, qoi, it & 1)
        want = P.encode(img, w, h, ch, it & 1, qoi)
        assert got == want, (it, w, h, first_difference(got, want))


@pytest.mark.parametrize("qoi", [0, 1])
def test_parallel_encoder_golden_vectors(emu, qoi):
    for v in golden("kat.json")["encode"]:
        if v["stream"] is None or v["channels"] < 3 or v["qoi"] != qoi:
            continue
        px = np.frombuffer(bytes.fromhex(v["pixels"]), dtype=np.uint8)
        ch = stored_channels(v["channels"])
        got = emu.encode(px, v["w"], v["h"], ch, qoi, v["colorspace"])
        assert got == bytes.fromhex(v["stream"]), v["name"]


@pytest.mark.parametrize("qoi", [0, 1])
def test_parallel_encoder_run_caps_across_tiles(emu, qoi):
    """Runs longer than a tile and longer than the run cap, cut at every offset class."""
    P = oracle.best()
    for n_run in (61, 62, 63, 511, 512, 513, 1023, 1024, 1025, 1536, 2047, 2048, 2049, 4099):
        for lead in (0, 1, 31, 33, 511):
            px = np.zeros((lead + n_run + 2, 4), dtype=np.uint8)
            px[:lead] = np.arange(lead * 4, dtype=np.uint32).reshape(lead, 4) % 251 + 1
            px[lead:lead + n_run] = (7, 8, 9, 255)
            px[lead + n_run:] = (1, 2, 3, 4)
            for tail in (0, 2):  # with / without a different pixel after the run
                img = px[: lead + n_run + tail]
                n = img.shape[0]
                got = emu.encode(img, n, 1, 4, qoi)
                want = P.encode(img, n, 1, 4, 0, qoi)
                assert got == want, (n_run, lead, tail, first_difference(got, want))


@pytest.mark.parametrize("qoi", [0, 1])
def test_parallel_encoder_batch_of_icons(emu, qoi):
    P = oracle.best()
    icons = synth.cfg3(24)
    got = emu.encode_batch(icons, 64, 64, 4, qoi)
    for i in range(24):
        want = P.encode(icons[i], 64, 64, 4, 0, qoi)
        assert got[i] == want, (i, first_difference(got[i], want))


def test_parallel_encoder_unaligned_rgb_and_partial_words(emu):
    P = oracle.best()
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5, 31, 32, 33, 43, 1023, 1025):
        for ch in (3, 4):
            buf = np.zeros(n * ch + 8, dtype=np.uint8)
            for shift in (0, 1, 2, 3):
                img = buf[shift: shift + n * ch]
                img[:] = random_image(rng, n, ch, 1).reshape(-1)
                got = emu.encode(img, n, 1, ch, 0)
                assert got == P.encode(img.copy(), n, 1, ch, 0, 0), (n, ch, shift)


# ---- parallel SQOA decoder -----------------------------------------------------------

@pytest.mark.parametrize("ch", [3, 4])
def test_parallel_sqoa_decoder_matches_oracle(emu, ch):
    P = oracle.best()
    rng = np.random.default_rng(300 + ch)
    for it in range(150):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        if it % 5 == 0:
            w, h = 1024, int(rng.integers(1, 9))
        if it % 9 == 0:
            w, h = 2048 + int(rng.integers(0, 3)), 5
        img = random_image(rng, w * h, ch, it % 4)
        if it % 13 == 0:
            img[:] = img[0]
        s = P.encode(img, w, h, ch, 0, 0)
        emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 777)
        for oc in (3, 4):
            got, st = emu.decode(s, w * h, ch, 0, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, w, h, oc, st)


def test_parallel_sqoa_decoder_golden_and_hostile_streams(emu):
    """Decoder-only behaviour through the parallel path: alpha suffix after runs / literals, an
    alpha byte at an op start, truncated bodies, over-long final runs, and REF ops (flagged and
    re-decoded by the serial code on the last thread block)."""
    P = oracle.best()
    n = 0
    for v in golden("kat.json")["decode"]:
        s = bytes.fromhex(v["stream"])
        w, h, hc, _cs, q = v["desc"]
        if q or hc < 3 or hc > 6 or len(s) < 22 or not s.startswith(b"Sqoa") or w * h == 0 or w * h > 10 ** 6:
            continue
        oc = v["channels"] or (3 if hc % 2 else 4)
        if oc not in (3, 4):
            continue
        got, st = emu.decode(s, w * h, hc, 0, oc)
        if v["pixels"] is None:
            assert st == -5, v["name"]
        else:
            assert st == 0 and got.tobytes() == bytes.fromhex(v["pixels"]), v["name"]
        n += 1
    assert n > 50
    rng = np.random.default_rng(3)
    decoded = 0
    for it in range(300):
        nb = int(rng.integers(22, 4000 if it % 7 == 0 else 200))
        s = bytearray(rng.integers(0, 256, nb, dtype=np.uint8).tobytes())
        w, h = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        s[0:4] = b"Sqoa"
        s[4:8] = w.to_bytes(4, "big")
        s[8:12] = h.to_bytes(4, "big")
        s[12] = int(rng.choice([3, 4, 5, 6]))
        s[13] = 0
        s[14] = 0x31
        mode = it % 4
        for k in range(15, nb):
            r = rng.random()
            if mode == 0 and r < 0.7:
                s[k] = int(rng.choice([0xFE, 0xFF, 0xFD, 0xC3, 0x85, 0x65, 0x70, 0x9A, 0xFC, 0xA0, 0x88]))
            elif mode == 1 and s[k] < 0x60:
                s[k] |= 0x80
            elif mode == 2 and r < 0.5:
                s[k] = int(rng.choice([0x60, 0x7F, 0x61, 0xFD, 0xFD, 0xC0]))
        s = bytes(s)
        emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 99)
        for oc in (3, 4):
            want, _ = P.decode(s, oc)
            got, st = emu.decode(s, w * h, s[12], 0, oc)
            if want is None:
                assert st == -5
            else:
                decoded += 1
                assert st == 0 and np.array_equal(got, want), (it, mode, oc)
    assert decoded > 300


def test_parallel_sqoa_decoder_batch_of_icons(emu):
    P = oracle.best()
    icons = synth.cfg3(20)
    streams = [P.encode(icons[i], 64, 64, 4, 0, 0) for i in range(20)]
    streams[7] = streams[7][:15] + bytes([0x00]) + streams[7][16:]  # a REF op: this image goes to the rescue path
    px, status = emu.decode_batch(streams, 64 * 64, 4, 0, 4)
    for i in range(20):
        want, _ = P.decode(streams[i], 4)
        if want is None:
            assert status[i] == -5
        else:
            assert status[i] == 0 and np.array_equal(px[i], want), i


# ---- one-launch QOI decoder for opaque streams (qoi_rows_kernels.cuh) ------------------------

def _photo(rng, w, h, ch, noise):
    """smooth gradient + noise: LUMA / DIFF / INDEX / RGB mix like cfg2"""
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255) // max(w, 1), (y * 255) // max(h, 1), (x + y) & 255], axis=-1)
    img = (base + rng.integers(-noise, noise + 1, (h, w, 3))) & 255
    if ch == 4:
        img = np.concatenate([img, np.full((h, w, 1), 255)], axis=-1)
    return img.astype(np.uint8).reshape(-1)


@pytest.mark.parametrize("ch", [3, 4])
def test_qoi_rows_kernel_decodes_opaque_streams_in_one_launch(emu, ch):
    """Streams without RGBA ops never leave the rows kernel: one launch, pixels identical to the reference's."""
    P = oracle.best()
    rng = np.random.default_rng(7100 + ch)
    emu.configure_qoi_rows(0)
    for it in range(40):
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 60))
        if it % 4 == 0:
            w, h = 1024 + int(rng.integers(0, 5)), int(rng.integers(4, 24))
        kind = it % 5
        if kind == 0:
            img = _photo(rng, w, h, ch, 3)
        elif kind == 1:
            img = _photo(rng, w, h, ch, 0)          # pure gradient: DIFF / LUMA chains, few literals
        elif kind == 2:
            pal = rng.integers(0, 256, (int(rng.integers(2, 90)), 3), dtype=np.uint8)   # INDEX- and run-heavy
            idx = np.repeat(rng.integers(0, len(pal), (w * h + 3) // 4), 4)[: w * h]
            idx = np.where(rng.random(w * h) < 0.5, idx, rng.integers(0, len(pal), w * h))
            img3 = pal[idx]
            img = (np.concatenate([img3, np.full((w * h, 1), 255, np.uint8)], axis=1) if ch == 4 else img3).reshape(-1)
        elif kind == 3:
            img = _photo(rng, w, h, ch, 40)         # literal-heavy
        else:
            img = _photo(rng, w, h, ch, 2)
            img.reshape(-1, ch)[rng.random(w * h) < 0.3] = img.reshape(-1, ch)[0]   # runs cut into everything
        s = P.encode(img, w, h, ch, 0, 1)
        emu.configure(int(rng.integers(1, 6)), int(rng.integers(0, 4)) * 131)
        for oc in (3, 4):
            before = emu.launch_count()
            got, st = emu.decode(s, w * h, ch, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, kind, w, h, oc, st)
            assert emu.launch_count() - before == 1, (it, kind, "fell back to the general pipeline")


def test_qoi_rows_kernel_long_runs_and_truncated_streams(emu):
    """Rows of RUN 62 ops (more pixels than the window holds), images that end inside a run, bodies that end
    early (the last pixel repeats, seqoia.h:726) and bodies longer than the image."""
    P = oracle.best()
    rng = np.random.default_rng(7200)
    emu.configure_qoi_rows(0)
    for it in range(24):
        w, h = 4096, int(rng.integers(1, 12))
        img = np.zeros((w * h, 3), np.uint8)
        img[:] = rng.integers(0, 256, 3)
        cuts = rng.integers(0, w * h, int(rng.integers(0, 40)))
        for c in cuts:
            img[c:] = rng.integers(0, 256, 3)
        s = bytearray(P.encode(img.reshape(-1), w, h, 3, 0, 1))
        if it % 3 == 1:   # drop ops from the end of the body
            cut = int(rng.integers(1, min(40, len(s) - 23)))
            s = s[: len(s) - 8 - cut] + s[-8:]
        if it % 3 == 2:   # claim a smaller image than the stream holds
            h = max(1, h - 1)
            s[8:12] = h.to_bytes(4, "big")
        s = bytes(s)
        emu.configure(int(rng.integers(1, 6)), it * 17)
        for oc in (3, 4):
            before = emu.launch_count()
            got, st = emu.decode(s, w * h, 3, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, oc, st)
            assert emu.launch_count() - before == 1, it


def test_qoi_general_pipeline_still_decodes_when_forced(emu):
    P = oracle.best()
    rng = np.random.default_rng(7300)
    emu.configure_qoi_rows(1)
    try:
        for it in range(10):
            w, h = int(rng.integers(1, 300)), int(rng.integers(1, 40))
            img = _photo(rng, w, h, 3, 3)
            s = P.encode(img, w, h, 3, 0, 1)
            before = emu.launch_count()
            got, st = emu.decode(s, w * h, 3, 1, 3)
            want, _ = P.decode(s, 3)
            assert st == 0 and np.array_equal(got, want), it
            assert emu.launch_count() - before > 1
    finally:
        emu.configure_qoi_rows(0)


def test_qoi_rows_kernel_hands_on_what_it_is_not_made_for(emu):
    """A read of a never-written slot other than 0 always goes to the general pipeline; RGBA ops and a read of the
    all-zero slot 0 only under a 3-channel header (under a 4-channel header the rows kernel tracks alpha)."""
    P = oracle.best()
    emu.configure_qoi_rows(0)
    end = bytes(7) + b"\x01"
    bodies = [
        bytes([0xFE, 10, 20, 30, 0x05, 0xFE, 1, 2, 3, 0x05, 0xC1]),        # INDEX into a never-written slot
        bytes([0xFF, 10, 20, 30, 40, 0xFE, 1, 2, 3, 0xC3]),                 # RGBA op
        bytes([0x00, 0xFE, 9, 9, 9, 0x00, 0xC2]),                           # slot 0 before anything wrote it
    ]
    for hc in (3, 4):
        hdr = b"qoif" + (6).to_bytes(4, "big") + (1).to_bytes(4, "big") + bytes([hc, 0])
        for i, body in enumerate(bodies):
            s = hdr + body + end
            for oc in (3, 4):
                before = emu.launch_count()
                got, st = emu.decode(s, 6, hc, 1, oc)
                want, _ = P.decode(s, oc)
                assert st == 0 and np.array_equal(got, want), (hc, i, oc)
                handed_on = emu.launch_count() - before > 1
                assert handed_on == (hc == 3 or i == 0), (hc, i, oc)


def _sprite(rng, w, h, kind):
    """RGBA content: transparent background, opaque and half-transparent shapes, antialiased edges"""
    img = np.zeros((h, w, 4), np.uint8)
    y, x = np.mgrid[0:h, 0:w]
    for _ in range(int(rng.integers(2, 7))):
        cx, cy, r = rng.integers(0, w), rng.integers(0, h), rng.integers(3, max(4, min(w, h) // 2 + 4))
        d = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
        inside = d < r
        col = rng.integers(0, 256, 3)
        if kind == 0:      # flat palette colours (INDEX / RUN heavy)
            img[inside, :3] = col
            img[inside, 3] = 255
        elif kind == 1:    # shaded, antialiased rim
            shade = (col[None, None, :] + (x[..., None] // 3)) & 255
            img[inside, :3] = shade[inside]
            img[inside, 3] = 255
            rim = (d >= r - 1.5) & inside
            img[rim, 3] = (255 * (r - d[rim]) / 1.5).astype(np.uint8)
        else:              # half-transparent noise
            img[inside, :3] = rng.integers(0, 256, (int(inside.sum()), 3))
            img[inside, 3] = rng.choice([40, 128, 255], int(inside.sum()))
    return img.reshape(-1)


def test_qoi_rows_kernel_tracks_alpha_under_a_4_channel_header(emu):
    """RGBA streams of the reference encoder: every pixel and alpha as the reference's, whichever path a stream takes;
    sprite-like content (few RGB literals after INDEX ops) must stay on the rows kernel."""
    P = oracle.best()
    rng = np.random.default_rng(7500)
    emu.configure_qoi_rows(0)
    stayed = {0: 0, 1: 0, 2: 0}
    for it in range(45):
        w, h = int(rng.integers(8, 300)), int(rng.integers(8, 90))
        if it % 5 == 0:
            w, h = 512, int(rng.integers(30, 70))
        kind = it % 3
        img = _sprite(rng, w, h, kind)
        s = P.encode(img, w, h, 4, 0, 1)
        emu.configure(int(rng.integers(1, 6)), int(rng.integers(0, 4)) * 71)
        for oc in (4, 3):
            before = emu.launch_count()
            got, st = emu.decode(s, w * h, 4, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, kind, w, h, oc, st)
            if oc == 4 and emu.launch_count() - before == 1:
                stayed[kind] += 1
    assert stayed[0] == 15 and stayed[1] >= 8, stayed


def test_qoi_rows_kernel_second_attempt_decodes_images_whose_alpha_guesses_fail(emu):
    """Icons where RGB literals follow INDEX ops that changed alpha: whichever stage decodes them (first attempt with
    the present guess model), every pixel is the reference's and the interpreter is not needed."""
    P = oracle.best()
    emu.configure_qoi_rows(0)
    for seed in (1054, 1058):
        w, h = 503, 130
        img = synth.image("icon", 503, 520, 4, seed=seed).reshape(520, 503 * 4)[200:330].reshape(-1).copy()
        s = P.encode(img, w, h, 4, 0, 1)
        emu.configure(3, seed)
        for oc in (4, 3):
            before = emu.launch_count()
            got, st = emu.decode(s, w * h, 4, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (seed, oc)
            assert emu.launch_count() - before == 1, (seed, oc, emu.launch_count() - before)


@pytest.mark.parametrize("whole_group", [0, 1])
def test_qoi_batch_mixes_opaque_rgba_and_hostile_streams(emu, whole_group):
    """One batch: opaque images stay on the rows kernel, images with RGBA ops or reads of never-written slots are
    flagged and decoded again by the general pipeline -- alone (their own image table) or with the whole group."""
    P = oracle.best()
    rng = np.random.default_rng(7400)
    w, h = 96, 40
    streams = []
    for i in range(14):
        kind = i % 4
        if kind == 0:
            img = _photo(rng, w, h, 4, 3)
        elif kind == 1:
            img = _photo(rng, w, h, 4, 2)
            img.reshape(-1, 4)[rng.random(w * h) < 0.2, 3] = 77          # alpha moves: RGBA ops
        elif kind == 2:
            img = _photo(rng, w, h, 4, 30)
        else:
            img = np.zeros(w * h * 4, np.uint8)                          # transparent black: INDEX 0 before any write
            img.reshape(-1, 4)[w * h // 2:] = (9, 9, 9, 255)
        streams.append(P.encode(img, w, h, 4, 0, 1))
    emu.configure_qoi_rows(0)
    emu.configure_qoi_fallback(whole_group)
    try:
        emu.configure(3, 5)
        px, status = emu.decode_batch(streams, w * h, 4, 1, 4)
    finally:
        emu.configure_qoi_fallback(0)
    for i in range(len(streams)):
        want, _ = P.decode(streams[i], 4)
        assert status[i] == 0 and np.array_equal(px[i], want), (i, status[i])


def _half_transparent_palette(rng, w, h):
    """alpha 128 everywhere, colours from a palette with fresh ones in between: RGB literals follow INDEX ops whose
    alpha is 128, not the 255 the rows kernel guesses"""
    pal = rng.integers(0, 256, (40, 3), dtype=np.uint8)
    img = np.zeros((w * h, 4), np.uint8)
    img[:, :3] = pal[rng.integers(0, 40, w * h)]
    fresh = rng.random(w * h) < 0.3
    img[fresh, :3] = rng.integers(0, 256, (int(fresh.sum()), 3))
    img[:, 3] = 128
    return img.reshape(-1)


def test_qoi_rows_kernel_chained_attempt_on_half_transparent_images(emu):
    P = oracle.best()
    rng = np.random.default_rng(7700)
    emu.configure_qoi_rows(0)
    for it in range(4):
        w, h = 300, 60 + 20 * it
        s = P.encode(_half_transparent_palette(rng, w, h), w, h, 4, 0, 1)
        emu.configure(1 + it, it)
        for oc in (4, 3):
            before = emu.qoi_stage_counts()
            got, st = emu.decode(s, w * h, 4, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, oc)
            after = emu.qoi_stage_counts()
            # the guesses of the rows kernel fail (alpha is 128, not 255), the general pipeline does not settle in its
            # three rounds, the chained attempt decodes the image; the interpreter is not needed
            assert tuple(a - b for a, b in zip(after, before)) == (1, 1, 0), (it, oc, before, after)


def test_qoi_batch_with_streams_for_every_attempt(emu):
    """One batch whose images end on different stages: opaque photos and an icon (first attempt of the rows kernel),
    a half-transparent palette image (general pipeline does not settle: chained rows attempt) and hand-made streams that
    read never-written slots (general pipeline) -- the later stages work on their own, smaller image tables."""
    P = oracle.best()
    rng = np.random.default_rng(7600)
    w, h = 503, 130
    n_px = w * h
    streams = []
    for i in range(3):
        streams.append(P.encode(_photo(rng, w, h, 4, 3), w, h, 4, 0, 1))
    icon = synth.image("icon", 503, 520, 4, seed=1054).reshape(520, 503 * 4)[200:330].reshape(-1).copy()
    streams.append(P.encode(icon, w, h, 4, 0, 1))
    streams.append(P.encode(_half_transparent_palette(rng, w, h), w, h, 4, 0, 1))
    hdr = b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([4, 0])
    end = bytes(7) + b"\x01"
    body = bytes([0xFE, 10, 20, 30, 0x05, 0xFE, 1, 2, 3, 0x05, 0xC1]) + bytes([0xFD]) * 900 + bytes([0x07, 0xFE, 5, 6, 7, 0x07])
    streams.append(hdr + body + end)                      # INDEX 5 / 7 before anything wrote those slots
    streams.append(P.encode(_photo(rng, w, h, 4, 25), w, h, 4, 0, 1))
    streams.append(hdr + bytes([0x21, 0xC5, 0x21, 0xFF, 1, 2, 3, 4, 0x21]) + end)
    emu.configure_qoi_rows(0)
    for whole_group in (0, 1):
        emu.configure_qoi_fallback(whole_group)
        try:
            emu.configure(4, 11 + whole_group)
            before = emu.launch_count()
            stages0 = emu.qoi_stage_counts()
            px, status = emu.decode_batch(streams, n_px, 4, 1, 4)
            launches = emu.launch_count() - before
            stages = tuple(a - b for a, b in zip(emu.qoi_stage_counts(), stages0))
        finally:
            emu.configure_qoi_fallback(0)
        # general pipeline, then the chained rows attempt; the hand-made streams may end on the interpreter
        assert launches > 4 and stages[:2] == (1, 1), (launches, stages)
        for i, s in enumerate(streams):
            want, _ = P.decode(s, 4)
            assert status[i] == 0 and np.array_equal(px[i], want), (whole_group, i, status[i])


def test_qoi_lane_tile_decodes_streams_without_alpha(emu):
    """The experimental lane-per-chunk tile for 3-channel QOI streams (qoi_lanes_kernels.cuh, SQOA_B200_QOI_LANES=1; off
    by default, measured slower than the rows tile): photo-like, flat, palette (INDEX
    chains across chunks and tiles), long runs, forced 4-channel output, unaligned output; same pixels as the reference
    and as the rows tile.  INDEX-heavy tiles (more own-chunk hits than the per-lane list holds) are handed to the rows
    tile; hostile streams (reads of never-written slots, an RGBA op) are flagged and end on the later stages."""
    P = oracle.best()
    rng = np.random.default_rng(9100)
    emu.configure_qoi_rows(0)
    emu.configure_qoi_lanes(0)  # on
    t0, b0 = emu.lanes_stats()
    for it in range(10):
        w, h = int(rng.integers(150, 900)), int(rng.integers(20, 70))
        kind = it % 5
        if kind == 0:
            img = _photo(rng, w, h, 3, 3)
        elif kind == 1:
            img = _photo(rng, w, h, 3, 40)
        elif kind == 2:  # palette: INDEX ops dominate
            pal = rng.integers(0, 256, (24, 3), dtype=np.uint8)
            img = pal[rng.integers(0, 24, w * h)].reshape(-1)
        elif kind == 3:  # long runs with photo stripes
            img = _photo(rng, w, h, 3, 5).reshape(h, w, 3)
            img[::3] = img[0, 0]
            img = img.reshape(-1)
        else:
            img = synth.image("mixed", w, h, 3, seed=300 + it, cell=(41, 9)).reshape(-1)
        s = P.encode(img, w, h, 3, 0, 1)
        for oc in (3, 4):
            want, _ = P.decode(s, oc)
            emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 313)
            got, st = emu.decode(s, w * h, 3, 1, oc)
            assert st == 0 and np.array_equal(got, want), (it, w, h, oc)
            emu.configure_qoi_lanes(1)  # off: the rows tile
            try:
                got2, st2 = emu.decode(s, w * h, 3, 1, oc)
            finally:
                emu.configure_qoi_lanes(0)
            assert st2 == 0 and np.array_equal(got2, want), ("rows tile", it, oc)
    t1, b1 = emu.lanes_stats()
    assert t1 - t0 > 50, (t0, t1)           # the lane tile did decode tiles
    assert b1 - b0 > 0, (b0, b1)            # ... and handed the INDEX-heavy ones to the rows tile
    # hostile: INDEX of a never-written slot, an RGBA op under the 3-channel header
    w, h = 400, 30
    hdr = b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([3, 0])
    end = bytes(7) + b"\x01"
    for body in (bytes([0xFE, 10, 20, 30, 0x05, 0x21, 0xC3]) + bytes([0xA0, 0x88] * 3000) + bytes([0x07, 0xFE, 5, 6, 7, 0x07]),
                 bytes([0xFE, 1, 2, 3, 0xFF, 9, 8, 7, 200, 0x6A]) + bytes([0xA2, 0x79] * 2500)):
        s = hdr + body + end
        for oc in (3, 4):
            want, _ = P.decode(s, oc)
            got, st = emu.decode(s, w * h, 3, 1, oc)
            assert st == 0 and np.array_equal(got, want), (len(body), oc)
    emu.configure_qoi_lanes(1)  # back to the library's default


def test_qoi_rows_tile_with_tiny_buffers():
    """The rows tile built with an op list of 64 entries, a 48-pixel window and 24 patches (tests/emu/Makefile, `tiny`):
    every tile is walked in many segments, flushes its window every row or two and overflows its patch list -- the
    paths a photo-like tile takes once in a while with the product's sizes.  Opaque, alpha-tracking, palette and icon
    streams, 3- and 4-byte output, batches; same pixels as the reference."""
    import subprocess

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "tiny"], check=True)
    tiny = Emu("libsqoa_emu_tiny.so")
    P = oracle.best()
    rng = np.random.default_rng(9300)
    tiny.configure_qoi_rows(0)
    streams = []
    for it in range(8):
        w, h = int(rng.integers(150, 700)), int(rng.integers(20, 60))
        ch = 3 + (it & 1)
        kind = it % 4
        if kind == 0:
            img = _photo(rng, w, h, ch, 3)
        elif kind == 1:
            img = synth.image("icon", w, h, ch, seed=400 + it).reshape(-1)
        elif kind == 2:
            img = synth.image("mixed", w, h, ch, seed=500 + it, cell=(29, 7)).reshape(-1)
        else:
            img = _photo(rng, w, h, ch, 2)
            if ch == 4:
                img.reshape(-1, 4)[rng.random(w * h) < 0.1, 3] = 200  # RGBA ops
        s = P.encode(img, w, h, ch, 0, 1)
        for oc in (3, 4):
            tiny.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 131)
            got, st = tiny.decode(s, w * h, ch, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, w, h, ch, oc)
        if ch == 4 and len(streams) < 4:
            streams.append((s, w * h))
    # icons in one batch (streams of one to a few tiles: decoded tile after tile)
    icons = [P.encode(synth.image("icon", 64, 64, 4, seed=700 + i).reshape(-1), 64, 64, 4, 0, 1) for i in range(12)]
    px, status = tiny.decode_batch(icons, 64 * 64, 4, 1, 4)
    for i, s in enumerate(icons):
        want, _ = P.decode(s, 4)
        assert status[i] == 0 and np.array_equal(px[i], want), i


def test_qoi_nowait_mode_decodes_every_kind_of_stream(emu):
    """sqoa_b200_ctx_set_qoi_nowait: all stages queued up front (rows, mark, chained retry grid, done, reset, interpreter),
    nothing read back in between; the same batch as above -- images that end on each stage -- and single images, then
    the default mode again on the same workspace (counters agree)."""
    P = oracle.best()
    rng = np.random.default_rng(7650)
    w, h = 503, 130
    n_px = w * h
    streams = [P.encode(_photo(rng, w, h, 4, 3), w, h, 4, 0, 1) for _ in range(3)]
    streams.append(P.encode(_half_transparent_palette(rng, w, h), w, h, 4, 0, 1))
    hdr = b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([4, 0])
    end = bytes(7) + b"\x01"
    body = bytes([0xFE, 10, 20, 30, 0x05, 0xFE, 1, 2, 3, 0x05, 0xC1]) + bytes([0xFD]) * 900 + bytes([0x07, 0xFE, 5, 6, 7, 0x07])
    streams.append(hdr + body + end)
    img = _photo(rng, w, h, 4, 2)
    img.reshape(-1, 4)[rng.random(n_px) < 0.2, 3] = 77
    streams.append(P.encode(img, w, h, 4, 0, 1))
    streams.append(hdr + bytes([0x21, 0xC5, 0x21, 0xFF, 1, 2, 3, 4, 0x21]) + end)
    emu.configure_qoi_rows(0)
    emu.configure_qoi_nowait(1)
    try:
        for rounds in range(2):
            emu.configure(3 + rounds, 17 * rounds)
            before = emu.launch_count()
            px, status = emu.decode_batch(streams, n_px, 4, 1, 4)
            assert emu.launch_count() - before == 6
            for i, s in enumerate(streams):
                want, _ = P.decode(s, 4)
                assert status[i] == 0 and np.array_equal(px[i], want), (rounds, i, status[i])
        for i, s in enumerate(streams):
            for oc in (3, 4):
                got, st = emu.decode(s, n_px, 4, 1, oc)
                want, _ = P.decode(s, oc)
                assert st == 0 and np.array_equal(got, want), (i, oc)
        opaque3 = P.encode(_photo(rng, w, h, 3, 4), w, h, 3, 0, 1)
        got, st = emu.decode(opaque3, n_px, 3, 1, 3)
        assert st == 0 and np.array_equal(got, P.decode(opaque3, 3)[0])
    finally:
        emu.configure_qoi_nowait(0)
    px, status = emu.decode_batch(streams, n_px, 4, 1, 4)
    for i, s in enumerate(streams):
        want, _ = P.decode(s, 4)
        assert status[i] == 0 and np.array_equal(px[i], want), ("default mode after", i, status[i])


# ---- parallel QOI decoder (scan / link / jump / verify / emit) ---------------------------

@pytest.mark.parametrize("ch", [3, 4])
def test_parallel_qoi_decoder_matches_oracle(emu, ch):
    P = oracle.best()
    rng = np.random.default_rng(500 + ch)
    for it in range(60):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        if it % 5 == 0:
            w, h = 1024, int(rng.integers(1, 9))
        if it % 9 == 0:
            w, h = 2048 + int(rng.integers(0, 3)), 5
        img = random_image(rng, w * h, ch, it % 4)
        if it % 13 == 0:
            img[:] = img[0]
        s = P.encode(img, w, h, ch, 0, 1)
        emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 777)
        for oc in (3, 4):
            got, st = emu.decode(s, w * h, ch, 1, oc)
            want, _ = P.decode(s, oc)
            assert st == 0 and np.array_equal(got, want), (it, w, h, oc, st)


def test_parallel_qoi_decoder_golden_and_hostile_streams(emu):
    """INDEX into never-written slots, a run as the first op (plants the start pixel in slot 53),
    RGB literals after INDEX ops with other alphas (the guess/verify loop has to iterate),
    truncated bodies -- all through the parallel pipeline."""
    P = oracle.best()
    n = 0
    for v in golden("kat.json")["decode"]:
        s = bytes.fromhex(v["stream"])
        w, h, hc, _cs, q = v["desc"]
        if not q or hc < 3 or hc > 6 or len(s) < 22 or v["pixels"] is None or w * h == 0 or w * h > 10 ** 6:
            continue
        oc = v["channels"] or (3 if hc % 2 else 4)
        if oc not in (3, 4):
            continue
        got, st = emu.decode(s, w * h, hc, 1, oc)
        assert st == 0 and got.tobytes() == bytes.fromhex(v["pixels"]), v["name"]
        n += 1
    assert n > 40
    rng = np.random.default_rng(5)
    decoded = 0
    for it in range(120):
        nb = int(rng.integers(22, 3000 if it % 7 == 0 else 200))
        s = bytearray(rng.integers(0, 256, nb, dtype=np.uint8).tobytes())
        w, h = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        s[0:4] = b"qoif"
        s[4:8] = w.to_bytes(4, "big")
        s[8:12] = h.to_bytes(4, "big")
        s[12] = int(rng.choice([3, 4, 5, 6]))
        s[13] = 0
        mode = it % 4
        for k in range(14, nb):
            r = rng.random()
            if mode == 0 and r < 0.7:
                s[k] = int(rng.choice([0xFE, 0xFF, 0xFD, 0xC3, 0x85, 0x65, 0x70, 0x9A, 0xFC, 0x05, 0x00, 0x35, 0x3F]))
            elif mode == 1 and r < 0.5:
                s[k] = int(rng.integers(0, 64))
            elif mode == 2 and r < 0.5:
                s[k] = int(rng.choice([0xFE, 0x00, 0x01, 0x35, 0xC0, 0xFF]))
        s = bytes(s)
        emu.configure(int(rng.integers(1, 5)), int(rng.integers(0, 3)) * 99)
        for oc in (3, 4):
            want, _ = P.decode(s, oc)
            if want is None:
                continue
            got, st = emu.decode(s, w * h, s[12], 1, oc)
            decoded += 1
            assert st == 0 and np.array_equal(got, want), (it, mode, oc)
    assert decoded > 150


def test_parallel_qoi_decoder_batch_of_icons(emu):
    P = oracle.best()
    icons = synth.cfg3(12)
    streams = [P.encode(icons[i], 64, 64, 4, 0, 1) for i in range(12)]
    px, status = emu.decode_batch(streams, 64 * 64, 4, 1, 4)
    for i in range(12):
        want, _ = P.decode(streams[i], 4)
        assert status[i] == 0 and np.array_equal(px[i], want), i


# ---- scanline shards of one image ----------------------------------------------

@pytest.mark.parametrize("qoi", [0, 1])
def test_sharded_encode_equals_whole_image_encode(emu, qoi):
    """Cut an image into shards at arbitrary pixel positions; summaries -> fold -> per-shard encode;
    the concatenated segments must be the reference's stream, byte for byte."""
    import seqoia_b200 as sb

    P = oracle.best()
    rng = np.random.default_rng(900 + qoi)
    for it in range(40):
        ch = int(rng.choice([3, 4]))
        w, h = int(rng.integers(8, 200)), int(rng.integers(2, 40))
        n = w * h
        img = random_image(rng, n, ch, it % 4)
        if it % 7 == 0:
            img[:] = img[0]                      # one run across every cut
        if it % 7 == 1:
            img[n // 3: 2 * n // 3] = img[n // 3]   # a whole shard inside a run
        n_shards = int(rng.integers(2, 6))
        cuts = sorted(set([0, n] + [int(x) for x in rng.integers(1, n, n_shards - 1)]))
        if it % 7 == 1:
            cuts = sorted(set(cuts + [n // 3 + 1, 2 * n // 3 - 1]))
        spans = list(zip(cuts[:-1], cuts[1:]))
        summaries = [emu.shard_summary(img[a:b], b - a, ch, qoi) for a, b in spans]
        out = b""
        for r, (a, b) in enumerate(spans):
            carry = sb.fold_carry(summaries, r, qoi)
            assert bytes(emu.fold_carry_device(summaries, r, qoi)) == bytes(carry), (it, r)  # device fold == host fold
            out += emu.encode(img[a:b], w, h, ch, qoi, it & 1, flags=4, carry=carry, n_px=b - a)
        want = P.encode(img, w, h, ch, it & 1, qoi)
        assert out == want, (it, w, h, ch, spans, first_difference(out, want))


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_sharded_sqoa_decode_equals_whole_decode(emu, ch):
    """SURVEY 8e, single image decode: byte ranges of one stream decoded separately, only the carry
    (entry offset, first pixel index, pixel before the shard) crosses shards."""
    P = oracle.best()
    rng = np.random.default_rng(900 + ch)
    for it in range(24):
        w, h = int(rng.integers(60, 300)), int(rng.integers(20, 60))
        img = random_image(rng, w * h, ch, it % 4)
        if it % 7 == 0:
            img[: (w * h) // 2] = img[0]   # a long run across shard boundaries
        s = P.encode(img, w, h, ch, 0, 0)
        want, _ = P.decode(s, ch)
        for n_shards in (2, 3):
            if (len(s) - 23) < 1920 * n_shards:
                continue
            emu.configure(int(rng.integers(1, 4)), int(rng.integers(0, 3)) * 4321)
            got = emu.decode_sharded(s, w * h, ch, ch, n_shards)
            assert np.array_equal(got, want), (it, w, h, n_shards)
            # the same with the carries folded by the device kernel and read from device memory (one-call path)
            got2, verdicts = emu.decode_sharded_device_fold(s, w * h, ch, ch, n_shards)
            assert all(v[0] == 0 for v in verdicts), verdicts
            assert np.array_equal(got2, want), (it, w, h, n_shards)
            assert verdicts[0][1] == 0 and sum(v[2] for v in verdicts) == w * h
            if it == 3:  # a pixel buffer that is too small: reported, nothing written past it
                _, tight = emu.decode_sharded_device_fold(s, w * h, ch, ch, n_shards, capacity_px=(w * h) // (n_shards * 4))
                assert any(v[0] == -3 for v in tight), tight


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_sharded_qoi_decode_equals_whole_decode(emu, ch):
    """A QOI stream in byte ranges, one after the other, each starting from the carry (64 slots, running pixel, entry,
    hash, pixel count) the range before it exported: pieces put together == the reference's pixels (seqoia.h:753-755,
    :785-787).  Ranges whose pixel buffer is too small say so; streams the optimistic attempt cannot decode are reported
    as not shardable by the range that found out and by every range after it."""
    P = oracle.best()
    rng = np.random.default_rng(8800 + ch)
    emu.configure_qoi_rows(0)
    for it in range(8):
        w, h = int(rng.integers(200, 700)), int(rng.integers(30, 90))
        kind = it % 4
        if kind == 0:
            img = _photo(rng, w, h, ch, 3)
        elif kind == 1:
            img = synth.image("mixed", w, h, ch, seed=100 + it, cell=(37, 11)).reshape(-1)
        elif kind == 2:
            img = synth.image("icon", w, h, ch, seed=200 + it).reshape(-1)
        else:
            img = _photo(rng, w, h, ch, 30)
        s = P.encode(img, w, h, ch, 0, 1)
        want, _ = P.decode(s, ch)
        for n_shards in (1, 2, 3, 5):
            emu.configure(int(rng.integers(1, 4)), int(rng.integers(0, 3)) * 977)
            got, verdicts, cuts = emu.qoi_decode_sharded(s, w * h, ch, ch, n_shards)
            assert all(v[0] == 0 for v in verdicts), (it, n_shards, verdicts)
            assert verdicts[0][1] == 0 and sum(v[2] for v in verdicts) == w * h, (it, n_shards, verdicts)
            for k in range(1, len(verdicts)):
                assert verdicts[k][1] == verdicts[k - 1][1] + verdicts[k - 1][2]
            assert np.array_equal(got, want), (it, w, h, n_shards)
        if it == 1:  # forced 3 <-> 4 channel output
            oc = 7 - ch
            got, verdicts, _ = emu.qoi_decode_sharded(s, w * h, ch, oc, 3)
            assert all(v[0] == 0 for v in verdicts) and np.array_equal(got, P.decode(s, oc)[0])
        if it == 2:  # a pixel buffer that is too small: reported with what the range needs, nothing written past it
            _, tight, _ = emu.qoi_decode_sharded(s, w * h, ch, ch, 3, capacity_px=(w * h) // 12)
            assert any(v[0] == -3 and v[2] > (w * h) // 12 for v in tight), tight
    if ch == 4:
        # half-transparent palette image: the alpha guesses of the optimistic attempt fail -> not shardable (-5) from the
        # range that notices on; the caller decodes such a stream on one GPU
        w, h = 300, 120
        s = P.encode(_half_transparent_palette(rng, w, h), w, h, 4, 0, 1)
        _, verdicts, _ = emu.qoi_decode_sharded(s, w * h, 4, 4, 3)
        assert any(v[0] == -5 for v in verdicts), verdicts
        first_bad = min(k for k, v in enumerate(verdicts) if v[0] == -5)
        assert all(v[0] == -5 for v in verdicts[first_bad:]), verdicts
        # and the ordinary decode of the same stream still works afterwards (flag counters back in step)
        got, st = emu.decode(s, w * h, 4, 1, 4)
        assert st == 0 and np.array_equal(got, P.decode(s, 4)[0])


# ---- one image in pieces (what the pipelined sqoa_encode / sqoa_decode launch) ---------------------------------
@pytest.mark.parametrize("qoi", [0, 1])
@pytest.mark.parametrize("ch", [3, 4])
def test_encode_in_pieces_matches_reference(emu, qoi, ch):
    cpu = oracle.best()
    w, h = 331, 97  # 8 tiles of 4096 pixels
    img = synth.image("mixed", w, h, ch, seed=21 + ch)
    want = cpu.encode(img, w, h, ch, 0, qoi)
    for piece in (1, 3):
        assert emu.encode_pieces(img, w, h, ch, qoi, piece) == want, (piece, qoi, ch)


@pytest.mark.parametrize("qoi", [0, 1])
@pytest.mark.parametrize("kind,ch", [("photo", 3), ("mixed", 4)])
def test_decode_in_pieces_matches_reference(emu, qoi, kind, ch):
    cpu = oracle.best()
    w, h = 200, 120
    img = synth.image(kind, w, h, ch, seed=5)
    s = cpu.encode(img, w, h, ch, 0, qoi)
    n_tiles = (len(s) - 22 - (0 if qoi else 1) + 1919) // 1920
    assert n_tiles >= 6
    for piece in (2, 5):
        px, st, prog = emu.decode_pieces(s, w * h, ch, qoi, ch, piece)
        if st == 0:  # (a QOI stream the optimistic rows kernel flags takes the plain path on the host)
            assert np.array_equal(px, img.reshape(-1)), (piece, qoi, kind)
            n_pieces = (n_tiles + piece - 1) // piece
            seen = [int(v) for v in prog[:n_pieces]]
            # pixels complete after every piece (the trailing run of a stream may overshoot the image, seqoia.h:640-642)
            assert seen == sorted(seen) and w * h <= seen[-1] <= w * h + 512
        else:
            assert qoi == 1


@pytest.mark.parametrize("qoi", [0, 1])
def test_shard_summary_scans_back_only_as_far_as_needed(emu, qoi):
    """The summary looks at the shard's last 262,144 pixels first and at the rest only when those do not settle the
    previous-pixel / run / index state: both outcomes must equal a summary of the whole shard."""
    rng = np.random.default_rng(7 + qoi)
    n = 300_000
    noisy = rng.integers(0, 256, (n, 4), dtype=np.uint8)             # every slot seen within the tail
    flat = noisy.copy()
    flat[20_000:] = flat[20_000]                                        # the tail is one long run: the scan must go on
    few = noisy.copy()
    few[30_000:] = np.array([[1, 2, 3, 255], [9, 8, 7, 255]], dtype=np.uint8)[rng.integers(0, 2, n - 30_000)]  # two colours only
    for img in (noisy, flat, few):
        s = emu.shard_summary(img, n, 4, qoi)
        px = img.view(np.uint32).reshape(-1)
        diff = np.nonzero(px[1:] != px[:-1])[0] + 1
        last = int(diff[-1]) if len(diff) else 0
        assert s.all_run == (0 if len(diff) else 1) and s.tail_run == n - 1 - last if len(diff) else s.tail_run == n - 1
        assert s.first_px == int(px[0]) and s.last_px == int(px[-1])
        if qoi:
            c = img[diff].astype(np.uint32)
            hsh = (c[:, 0] * 3 + c[:, 1] * 5 + c[:, 2] * 7 + c[:, 3] * 11) & 63
            for slot in range(64):
                at = diff[hsh == slot]
                have = (s.slot_valid[slot >> 5] >> (slot & 31)) & 1
                assert have == (1 if len(at) else 0), slot
                if len(at):
                    assert s.slot_px[slot] == int(px[at[-1]]), slot
