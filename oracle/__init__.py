"""oracle -- TEST INFRASTRUCTURE ONLY (parity checker for the B200 codec).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``seqoia_b200/`` does.

Two CPU codecs with one Python face:

* ``port()``      -- ``liboracle.so``: the restatement in ``oracle/sqoa_oracle.c``.
* ``reference()`` -- ``_ref/libsqoa_ref.so``: the unmodified reference header
  compiled where it lies (``make -C oracle ref``); ``None`` when it has not been
  built (the GPU box uses the prebuilt file that travels with the snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Desc(C.Structure):
    """Same layout as ``sqoa_desc`` (seqoia.h:318-324)."""

    _fields_ = [
        ("width", C.c_uint),
        ("height", C.c_uint),
        ("channels", C.c_ubyte),
        ("colorspace", C.c_ubyte),
        ("qoi_compat", C.c_ubyte),
    ]


def build(with_reference: bool = True) -> None:
    """Compile the restatement (always) and the reference shim (when its sources exist)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if with_reference:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class CpuCodec:
    """A CPU codec with the reference's call shape (malloc-returning encode / decode)."""

    def __init__(self, path: str, enc: str, dec: str, free: str, kind: str):
        self.kind = kind
        self.path = path
        self.lib = C.CDLL(path)
        self._enc = getattr(self.lib, enc)
        self._enc.restype = C.c_void_p
        self._enc.argtypes = [C.c_void_p, C.POINTER(Desc), C.POINTER(C.c_int)]
        self._dec = getattr(self.lib, dec)
        self._dec.restype = C.c_void_p
        self._dec.argtypes = [C.c_void_p, C.c_int, C.POINTER(Desc), C.c_int]
        self._free = getattr(self.lib, free)
        self._free.restype = None
        self._free.argtypes = [C.c_void_p]

    # function pointers for the C timing driver (oracle/cpu_batch.c)
    @property
    def enc_ptr(self):
        return C.cast(self._enc, C.c_void_p)

    @property
    def dec_ptr(self):
        return C.cast(self._dec, C.c_void_p)

    def encode(self, pixels, width: int, height: int, channels: int, colorspace: int = 0,
               qoi: int = 0) -> Optional[bytes]:
        buf = np.ascontiguousarray(np.frombuffer(pixels, dtype=np.uint8) if not isinstance(pixels, np.ndarray)
                                   else pixels.reshape(-1).view(np.uint8))
        d = Desc(width, height, channels, colorspace, qoi)
        n = C.c_int(0)
        p = self._enc(buf.ctypes.data_as(C.c_void_p), C.byref(d), C.byref(n))
        if not p:
            return None
        out = C.string_at(p, n.value)
        self._free(p)
        return out

    def decode(self, stream, channels: int = 0, size: Optional[int] = None) -> Tuple[Optional[np.ndarray], Desc]:
        raw = np.frombuffer(bytes(stream), dtype=np.uint8)
        n = len(raw) if size is None else size
        # 64 zero bytes of slack so that a hostile stream can never make a CPU
        # codec read outside the buffer
        padded = np.zeros(len(raw) + 64, dtype=np.uint8)
        padded[: len(raw)] = raw
        d = Desc()
        p = self._dec(padded.ctypes.data_as(C.c_void_p), n, C.byref(d), channels)
        if not p:
            return None, d
        colour = 1 if d.channels < 3 else 3
        ch = channels if channels != 0 else colour + (1 if d.channels % 2 == 0 else 0)
        nbytes = d.width * d.height * ch
        out = np.frombuffer(C.string_at(p, nbytes), dtype=np.uint8).copy()
        self._free(p)
        return out, d


_port: Optional[CpuCodec] = None
_ref: Optional[CpuCodec] = None


def port() -> CpuCodec:
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(with_reference=False)
        _port = CpuCodec(path, "oracle_encode_alloc", "oracle_decode_alloc", "oracle_free", "port")
    return _port


def reference() -> Optional[CpuCodec]:
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libsqoa_ref.so")
        if not os.path.exists(path):
            return None
        _ref = CpuCodec(path, "ref_sqoa_encode", "ref_sqoa_decode", "ref_free", "reference")
    return _ref


def best() -> CpuCodec:
    """The compiled reference when present, else the restatement."""
    return reference() or port()


def timing_driver():
    """ctypes handle of oracle/cpu_batch.c (lives in liboracle.so)."""
    lib = port().lib
    lib.cb_max_threads.restype = C.c_int
    lib.cb_time_encode.restype = C.c_double
    lib.cb_time_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    lib.cb_time_decode.restype = C.c_double
    lib.cb_time_decode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_int,
                                   C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    return lib
