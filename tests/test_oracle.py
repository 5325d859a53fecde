"""The oracle restatement (oracle/sqoa_oracle.c) is pinned to the reference:
against the committed golden vectors (generated from the compiled reference by
oracle/make_golden.py) and, when oracle/_ref is present, against the compiled
reference itself on random images and random / hostile streams."""
import hashlib

import numpy as np
import pytest

import oracle
from seqoia_b200 import synth
from util import golden, random_image, stored_channels


def test_port_matches_golden_encode_vectors():
    P = oracle.port()
    kat = golden("kat.json")
    assert len(kat["encode"]) > 100
    for v in kat["encode"]:
        px = np.frombuffer(bytes.fromhex(v["pixels"]), dtype=np.uint8)
        got = P.encode(px, v["w"], v["h"], v["channels"], v["colorspace"], v["qoi"])
        want = None if v["stream"] is None else bytes.fromhex(v["stream"])
        assert got == want, v["name"]


def test_port_matches_golden_decode_vectors():
    P = oracle.port()
    kat = golden("kat.json")
    assert len(kat["decode"]) > 100
    for v in kat["decode"]:
        px, d = P.decode(bytes.fromhex(v["stream"]), v["channels"])
        want = None if v["pixels"] is None else bytes.fromhex(v["pixels"])
        assert (None if px is None else px.tobytes()) == want, v["name"]
        if want is not None:
            assert [d.width, d.height, d.channels, d.colorspace, d.qoi_compat] == v["desc"], v["name"]


def test_survey_appendix_c_bytes():
    """The two streams printed in SURVEY.md appendix C, literally."""
    P = oracle.port()
    px = np.array([10, 20, 30, 255, 10, 20, 30, 255, 11, 21, 31, 255, 200, 100, 50, 128], dtype=np.uint8)
    sq = "53716f61" "00000004" "00000001" "04" "00" "31" "fe0a141e" "c0" "a188" "ffc8643280" "0000000000000001"
    qf = "716f6966" "00000004" "00000001" "04" "00" "fe0a141e" "c0" "7f" "ffc8643280" "0000000000000001"
    assert P.encode(px, 4, 1, 4, 0, 0).hex() == sq
    assert P.encode(px, 4, 1, 4, 0, 1).hex() == qf


@pytest.mark.parametrize("name", ["cfg1_1920x1080_rgba", "cfg2_3840x2160_rgb"])
@pytest.mark.parametrize("qoi", [0, 1])
def test_port_matches_reference_digests_at_full_size(name, qoi):
    """BASELINE.json configs 0 and 1 at full size: same bytes as the reference produced."""
    P = oracle.port()
    dig = golden("digests.json")["digests"][f"{name}_q{qoi}"]
    img = synth.cfg1() if name.startswith("cfg1") else synth.cfg2()
    assert hashlib.sha256(img.tobytes()).hexdigest() == dig["pixels_sha256"], "generator drifted"
    s = P.encode(img, dig["w"], dig["h"], dig["channels"], 0, qoi)
    assert len(s) == dig["stream_len"]
    assert hashlib.sha256(s).hexdigest() == dig["stream_sha256"]
    back, d = P.decode(s, 0)
    assert np.array_equal(back, img.reshape(-1))


def _need_ref():
    R = oracle.reference()
    if R is None:
        pytest.skip("oracle/_ref/libsqoa_ref.so not built here (no /root/reference)")
    return R


def test_port_matches_compiled_reference_on_random_images():
    P, R = oracle.port(), _need_ref()
    rng = np.random.default_rng(11)
    for it in range(600):
        ch = int(rng.integers(1, 7))
        qoi = int(rng.integers(0, 2))
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 30))
        if it % 40 == 0:
            w, h = 1300, 2
        img = random_image(rng, w * h, stored_channels(ch), it % 4)
        a, b = P.encode(img, w, h, ch, it & 1, qoi), R.encode(img, w, h, ch, it & 1, qoi)
        assert a == b, (it, w, h, ch, qoi)
        if a is None:
            continue
        for oc in (0, 1, 2, 3, 4):
            pa, da = P.decode(a, oc)
            pb, db = R.decode(a, oc)
            assert np.array_equal(pa, pb), (it, oc)
            assert (da.width, da.height, da.channels, da.colorspace, da.qoi_compat) == (
                db.width, db.height, db.channels, db.colorspace, db.qoi_compat)


def test_port_matches_compiled_reference_on_hostile_streams():
    """sqoafuzz.c's intent: arbitrary bytes, incl. REF ops (seqoia.h:729-738)."""
    P, R = oracle.port(), _need_ref()
    rng = np.random.default_rng(12)
    decoded = 0
    for it in range(2500):
        n = int(rng.integers(22, 160))
        s = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        s[0:4] = b"Sqoa" if it % 2 == 0 else b"qoif"
        s[4:8] = int(rng.integers(1, 20)).to_bytes(4, "big")
        s[8:12] = int(rng.integers(1, 20)).to_bytes(4, "big")
        s[12] = int(rng.integers(1, 7))
        s[13] = int(rng.integers(0, 2))
        if it % 2 == 0:
            s[14] = 0x31
        if it % 3 == 0:
            for k in range(15, n):
                if rng.random() < 0.5:
                    s[k] = int(rng.choice([0xFE, 0xFF, 0xFD, 0xC3, 0x85, 0x65, 0x70, 0x9A, 0x41, 0x05]))
        for oc in (0, 3, 4, 1, 2):
            pa, da = P.decode(bytes(s), oc)
            pb, db = R.decode(bytes(s), oc)
            assert (pa is None) == (pb is None), (it, oc, bytes(s).hex())
            if pa is not None:
                decoded += 1
                assert np.array_equal(pa, pb), (it, oc, bytes(s).hex())
    assert decoded > 2000
