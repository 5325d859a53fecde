"""The drop-in boundary: libsqoa_b200.so loads, exports every symbol that
include/sqoa_b200.h declares, keeps the reference's struct layout, rejects the
arguments the reference rejects -- and refuses to work without a GPU instead of
falling back to a CPU codec.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import seqoia_b200 as sb
from util import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sqoa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(sqoa_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    L = sb.lib()
    names = declared_functions()
    assert {"sqoa_encode", "sqoa_decode", "sqoa_read", "sqoa_write"} <= set(names)
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sqoa_b200.h but not exported"


def test_desc_layout_matches_reference():
    # seqoia.h:318-324: two unsigned + three unsigned char, sizeof 12 on x86-64
    assert C.sizeof(sb.Desc) == 12
    assert sb.Desc.width.offset == 0 and sb.Desc.height.offset == 4
    assert sb.Desc.channels.offset == 8 and sb.Desc.colorspace.offset == 9 and sb.Desc.qoi_compat.offset == 10
    assert C.sizeof(sb.Item) == 32
    assert C.sizeof(sb.ShardSummary) == 80 * 4
    assert C.sizeof(sb.Carry) == 72 * 4


def test_header_compiles_as_c99_and_cxx():
    for comp, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
        subprocess.run([comp, std, "-Wall", "-Werror", "-fsyntax-only", "-x", lang,
                        os.path.join(ROOT, "include", "sqoa_b200.h")], check=True)


def test_encode_rejects_what_the_reference_rejects():
    px = np.zeros(64, dtype=np.uint8)
    assert sb.encode(px, 0, 1, 4) is None            # zero dimension   seqoia.h:467
    assert sb.encode(px, 1, 0, 4) is None
    assert sb.encode(px, 1, 1, 0) is None            # channels         seqoia.h:468
    assert sb.encode(px, 1, 1, 7) is None
    assert sb.encode(px, 1, 1, 4, colorspace=2) is None  # colorspace   seqoia.h:469
    assert sb.encode(px, 20000, 20000, 4) is None    # size cap         seqoia.h:470
    assert sb.encode(px, 2, 1, 1, qoi=1) is None     # mono + QOI       seqoia.h:477-480
    assert sb.encode(px, 2, 1, 2, qoi=1) is None
    L = sb.lib()
    d = sb.Desc(1, 1, 4, 0, 0)
    n = C.c_int(0)
    assert not L.sqoa_encode(None, C.byref(d), C.byref(n))
    assert not L.sqoa_encode(px.ctypes.data, None, C.byref(n))
    assert not L.sqoa_encode(px.ctypes.data, C.byref(d), None)


def test_decode_rejects_and_fills_desc_like_the_reference():
    hdr = b"Sqoa" + (3).to_bytes(4, "big") + (2).to_bytes(4, "big") + bytes([4, 1, 0x31])
    body = hdr + bytes([0xC0]) + bytes(7) + b"\x01"
    px, d = sb.decode(body[:21])                     # size < 22        seqoia.h:665
    assert px is None and d.width == 0
    px, d = sb.decode(body, channels=5)              # channels > 4     seqoia.h:664
    assert px is None and d.width == 0
    bad = b"Xqoa" + body[4:]
    px, d = sb.decode(bad)                           # magic: desc is filled first (seqoia.h:673-677)
    assert px is None and (d.width, d.height, d.channels, d.colorspace, d.qoi_compat) == (3, 2, 4, 1, 0)
    qoif = b"qoif" + body[4:]
    px, d = sb.decode(qoif)                          # qoif + start byte seqoia.h:684
    assert px is None and d.qoi_compat == 0
    rc, d, n = sb.probe(body[:15], len(body), 0)
    assert rc == sb.OK and n == 3 * 2 * 4
    rc, d, n = sb.probe(body[:15], len(body), 3)
    assert rc == sb.OK and n == 3 * 2 * 3
    no_start = b"Sqoa" + body[4:14] + bytes([0xC0])
    rc, d, n = sb.probe(no_start, 40, 0)             # Sqoa magic without start byte decodes as QOI
    assert rc == sb.OK and d.qoi_compat == 1


def test_max_stream_size():
    assert sb.max_stream_size(4, 1, 4) == 4 * 5 + 23
    assert sb.max_stream_size(4, 1, 3) == 4 * 4 + 23
    assert sb.max_stream_size(2, 1, 5) == 2 * 4 + 23   # BGR is stored as 3 channels
    assert sb.max_stream_size(2, 1, 2) == 2 * 3 + 23


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    px = np.arange(16, dtype=np.uint8)
    assert sb.encode(px, 2, 2, 4) is None
    with pytest.raises(sb.SqoaError) as e:
        sb.Context()
    assert "no CPU fallback" in str(e.value) or "sm_100" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """Nothing under seqoia_b200/ may import, include or link oracle/ or the emulator."""
    pkg = os.path.join(ROOT, "seqoia_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            for banned in ("liboracle", "import oracle", "from oracle", "oracle/", "libsqoa_emu", "emu_api"):
                assert banned not in text, f"{f} mentions {banned}"
    out = subprocess.run(["ldd", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emu" not in out
