"""The C command-line tools (tools/sqoabench_b200.c, tools/sqoaconv_b200.c): plain C programs linked against
libsqoa_b200.so through include/sqoa_b200.h -- the "real C caller" side of the drop-in boundary.
CPU part: they build and link.  GPU part: the harness runs the reference's round-trip verification
(sqoabench.c:446-455) and a byte comparison with the compiled reference on BASELINE cfg1; the converter round-trips
.raw -> .sqoa -> .qoi -> .raw (sqoaconv.c:38-100)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")
REF = os.path.join(ROOT, "oracle", "_ref", "libsqoa_ref.so")


@pytest.fixture(scope="module")
def tools():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    return BIN


def test_tools_build_and_link(tools):
    for name in ("sqoabench_b200", "sqoaconv_b200"):
        exe = os.path.join(tools, name)
        assert os.access(exe, os.X_OK)
        libs = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libsqoa_b200.so" in libs and "not found" not in libs, libs
    # usage text, no GPU needed
    out = subprocess.run([os.path.join(tools, "sqoaconv_b200")], capture_output=True, text=True)
    assert out.returncode == 1 and "Usage" in out.stdout


@pytest.mark.gpu
def test_sqoabench_cfg1_round_trip_and_reference_parity(tools):
    cmd = [os.path.join(tools, "sqoabench_b200"), "2", "--synth", "cfg1"]
    if os.path.exists(REF):
        cmd += ["--reference", REF]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
    rows = [ln.split()[0] for ln in out.stdout.splitlines() if ln and ln.split()[0] in
            ("sqoa", "qoi", "sqoa-dev", "qoi-dev", "ref-sqoa", "ref-qoi")]
    assert {"sqoa", "qoi", "sqoa-dev", "qoi-dev"} <= set(rows), out.stdout
    if os.path.exists(REF):
        assert {"ref-sqoa", "ref-qoi"} <= set(rows), out.stdout


@pytest.mark.gpu
def test_sqoabench_flags_and_directory(tools, tmp_path):
    from seqoia_b200 import synth

    d = tmp_path / "images"
    (d / "sub").mkdir(parents=True)
    synth.image("icon", 64, 64, 4, seed=5).tofile(d / "icon.64x64x4.raw")
    synth.image("photo", 200, 100, 3, seed=6).tofile(d / "sub" / "photo.200x100x3.raw")
    exe = os.path.join(tools, "sqoabench_b200")
    out = subprocess.run([exe, "1", str(d), "--onlytotals", "--nowarmup"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "Grand total (2 images" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([exe, "1", str(d), "--norecurse", "--noencode", "--noverify"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "all checks passed (1 image(s))" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_sqoaconv_round_trip(tools, tmp_path):
    import oracle
    from seqoia_b200 import synth

    exe = os.path.join(tools, "sqoaconv_b200")
    img = synth.image("mixed", 333, 77, 4, seed=9)
    raw = tmp_path / "in.333x77x4.raw"
    img.tofile(raw)
    run = lambda a, b: subprocess.run([exe, str(a), str(b)], capture_output=True, text=True, timeout=120)
    assert run(raw, tmp_path / "a.sqoa").returncode == 0
    assert run(tmp_path / "a.sqoa", tmp_path / "b.qoi").returncode == 0
    assert run(tmp_path / "b.qoi", tmp_path / "c.sqoa").returncode == 0
    assert run(tmp_path / "c.sqoa", tmp_path / "out.raw").returncode == 0
    assert np.array_equal(np.fromfile(tmp_path / "out.raw", dtype=np.uint8), img.reshape(-1))
    # the files are the reference's streams, byte for byte
    cpu = oracle.best()
    assert (tmp_path / "a.sqoa").read_bytes() == cpu.encode(img, 333, 77, 4, 0, 0)
    assert (tmp_path / "b.qoi").read_bytes() == cpu.encode(img, 333, 77, 4, 0, 1)
    assert (tmp_path / "c.sqoa").read_bytes() == (tmp_path / "a.sqoa").read_bytes()
    assert run(tmp_path / "missing.qoi", tmp_path / "x.sqoa").returncode == 1


@pytest.mark.gpu
def test_sqoaconv_many_files_in_one_call(tools, tmp_path):
    """sqoaconv_b200 --many: a C caller of sqoa_b200_read_many / sqoa_b200_write_many (the file path with
    many-file batching): every output is the reference's stream of the decoded input, byte for byte; an input that does
    not decode is reported and the others are still converted."""
    import oracle
    from seqoia_b200 import synth

    cpu = oracle.best()
    exe = os.path.join(tools, "sqoaconv_b200")
    src = tmp_path / "in"
    dst = tmp_path / "out"
    src.mkdir()
    dst.mkdir()
    shapes = [(64, 64, 4, "icon", 0), (300, 200, 3, "photo", 1), (257, 129, 4, "mixed", 1), (640, 480, 3, "screen", 0)]
    imgs = {}
    for k, (w, h, c, kind, q) in enumerate(shapes):
        img = synth.image(kind, w, h, c, seed=20 + k).reshape(-1)
        name = f"f{k}.{'qoi' if q else 'sqoa'}"
        (src / name).write_bytes(cpu.encode(img, w, h, c, 0, q))
        imgs[f"f{k}"] = (img, w, h, c)
    files = sorted(str(p) for p in src.iterdir())
    out = subprocess.run([exe, "--many", ".qoi", str(dst)] + files, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "4 of 4 files decoded, 4 written" in out.stdout, out.stdout + out.stderr
    for stem, (img, w, h, c) in imgs.items():
        assert (dst / f"{stem}.qoi").read_bytes() == cpu.encode(img, w, h, c, 0, 1), stem
    (src / "broken.sqoa").write_bytes(b"Sqoa" + bytes(30))
    out = subprocess.run([exe, "--many", ".sqoa", str(dst)] + files + [str(src / "broken.sqoa")], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 1 and "4 of 5 files decoded, 4 written" in out.stdout, out.stdout + out.stderr
    for stem, (img, w, h, c) in imgs.items():
        assert (dst / f"{stem}.sqoa").read_bytes() == cpu.encode(img, w, h, c, 0, 0), stem
