"""Deterministic synthetic images for the BASELINE.json configs.

Thin ctypes face of ``libsqoa_synth.so`` (``csrc/synth.c``).  Every pixel is a
pure function of (recipe, seed, x, y), so the GPU box regenerates the very same
bytes the golden digests in ``tests/golden`` were made from.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"mixed": 0, "photo": 1, "icon": 2, "screen": 3}
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsqoa_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(path)
        _lib.sqoa_synth_image.restype = C.c_int
        _lib.sqoa_synth_image.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint64, C.c_uint, C.c_uint,
                                          C.c_void_p, C.c_int]
        _lib.sqoa_synth_batch.restype = C.c_int
        _lib.sqoa_synth_batch.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint64, C.c_size_t,
                                          C.c_void_p, C.c_int]
        _lib.sqoa_synth_rows.restype = C.c_int
        _lib.sqoa_synth_rows.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_uint64, C.c_uint, C.c_uint, C.c_uint,
                                         C.c_uint, C.c_void_p, C.c_int]
    return _lib


def rows(kind: str, width: int, height: int, channels: int, y0: int, y1: int, seed: int = 42, cell=(0, 0),
         threads: int = 0) -> np.ndarray:
    """Rows ``[y0, y1)`` of :func:`image` (a scanline shard), shape ``(y1-y0, width, channels)``."""
    lib = _load()
    out = np.empty((y1 - y0, width, channels), dtype=np.uint8)
    rc = lib.sqoa_synth_rows(KINDS[kind], width, height, channels, seed, cell[0], cell[1], y0, y1,
                             out.ctypes.data_as(C.c_void_p), threads)
    if rc != 0:
        raise ValueError("bad synthetic row-range request")
    return out


def image(kind: str, width: int, height: int, channels: int, seed: int = 42, cell=(0, 0), out=None,
          threads: int = 0) -> np.ndarray:
    """One ``height x width x channels`` uint8 image of recipe ``kind``."""
    lib = _load()
    if out is None:
        out = np.empty((height, width, channels), dtype=np.uint8)
    assert out.nbytes == width * height * channels and out.flags["C_CONTIGUOUS"]
    rc = lib.sqoa_synth_image(KINDS[kind], width, height, channels, seed, cell[0], cell[1],
                              out.ctypes.data_as(C.c_void_p), threads)
    if rc != 0:
        raise ValueError("bad synthetic image request")
    return out


def batch(kind: str, n: int, width: int, height: int, channels: int, seed0: int = 0, out=None,
          threads: int = 0) -> np.ndarray:
    """``n`` images, image ``i`` seeded with ``seed0 + i``, shape ``(n, h, w, c)``."""
    lib = _load()
    if out is None:
        out = np.empty((n, height, width, channels), dtype=np.uint8)
    stride = width * height * channels
    assert out.nbytes == n * stride and out.flags["C_CONTIGUOUS"]
    rc = lib.sqoa_synth_batch(KINDS[kind], n, width, height, channels, seed0, stride,
                              out.ctypes.data_as(C.c_void_p), threads)
    if rc != 0:
        raise ValueError("bad synthetic batch request")
    return out


# The five BASELINE.json configs ------------------------------------------------

def cfg1(width: int = 1920, height: int = 1080) -> np.ndarray:
    """single 1920x1080 RGBA image: gradients + noise + flat regions."""
    return image("mixed", width, height, 4, seed=42)


def cfg2(width: int = 3840, height: int = 2160, channels: int = 3) -> np.ndarray:
    """3840x2160 RGB photo-like image."""
    return image("photo", width, height, channels, seed=42)


def cfg3(n: int = 100_000, first: int = 0) -> np.ndarray:
    """icons ``first .. first+n`` of the 100k 64x64 RGBA batch (image i seeded with i)."""
    return batch("icon", n, 64, 64, 4, seed0=first)


def cfg4(width: int = 20000, height: int = 19999, out=None) -> np.ndarray:
    """the largest 20000-wide RGBA image under the 400 Mpx cap; cfg1's recipe with scaled cells."""
    scale = max(1, width // 1920)
    return image("mixed", width, height, 4, seed=42, cell=(97 * scale, 53 * scale), out=out)


def cfg4_rows(y0: int, y1: int, width: int = 20000, height: int = 19999) -> np.ndarray:
    """scanlines ``[y0, y1)`` of :func:`cfg4`: what one GPU holds when the image is sharded."""
    scale = max(1, width // 1920)
    return rows("mixed", width, height, 4, y0, y1, seed=42, cell=(97 * scale, 53 * scale))


# (directory, count, mean Mpx, recipe, channels) of the qoi benchmark suite, derived in SURVEY.md 8d
CFG5_MIX = [
    ("icon_64", 217, 0.004, "icon", 4), ("icon_512", 213, 0.262, "icon", 4),
    ("textures_pk", 996, 0.045, "photo", 4), ("textures_pk01", 113, 0.130, "photo", 4),
    ("textures_pk02", 236, 0.303, "photo", 4), ("screenshot_game", 618, 0.633, "screen", 3),
    ("screenshot_web", 14, 8.124, "screen", 4), ("pngimg", 187, 1.808, "icon", 4),
    ("textures_plants", 60, 1.065, "photo", 4), ("textures_photo", 20, 1.049, "photo", 3),
    ("photo_kodak", 24, 0.393, "photo", 3), ("photo_tecnick", 100, 1.438, "photo", 3),
    ("photo_wikipedia", 49, 1.084, "photo", 3),
]


def cfg5_shapes(scale: float = 1.0):
    """(kind, width, height, channels, seed) for every image of the mixed corpus."""
    shapes = []
    seed = 1000
    for _name, count, mpx, kind, ch in CFG5_MIX:
        n = max(1, int(round(count * scale)))
        side = max(8, int(round((mpx * 1e6) ** 0.5)))
        for k in range(n):
            w = max(8, side + (k % 7) * 3 - 9)
            h = max(8, int(mpx * 1e6 / w))
            shapes.append((kind, w, h, ch, seed))
            seed += 1
    return shapes
