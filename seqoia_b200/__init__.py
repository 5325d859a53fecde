"""seqoia_b200 -- Python face of ``libsqoa_b200.so`` (the C ABI in ``include/sqoa_b200.h``).

The product is the shared library; this module only binds it with ctypes so the
tests and ``bench.py`` can call exactly what a C program would call:

* :func:`encode` / :func:`decode` / :func:`write` / :func:`read` -- the reference's
  four entry points (``seqoia.h:336-374``) on host memory.
* :class:`Context` -- the device-resident extension surface (single image, batch,
  shards) on raw device pointers; PyTorch is used by callers only to own device
  memory and streams.

Loading fails loudly when the library has not been built, and every call fails
(``None`` / exception) when no B200 is present: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SQOA_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("SQOA_B200_LIB") or os.path.join(_HERE, "libsqoa_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

OK, E_ARG, E_CUDA, E_CAPACITY, E_NOGPU, E_STREAM = 0, -1, -2, -3, -4, -5
PATH_AUTO, PATH_PARALLEL, PATH_SERIAL = 0, 1, 2


class Desc(C.Structure):
    """``sqoa_desc`` (replaces seqoia.h:318-324)."""

    _fields_ = [
        ("width", C.c_uint),
        ("height", C.c_uint),
        ("channels", C.c_ubyte),
        ("colorspace", C.c_ubyte),
        ("qoi_compat", C.c_ubyte),
    ]


class Item(C.Structure):
    """``sqoa_b200_item``."""

    _fields_ = [
        ("in_offset", C.c_ulonglong),
        ("out_offset", C.c_ulonglong),
        ("width", C.c_uint),
        ("height", C.c_uint),
        ("size", C.c_uint),
        ("channels", C.c_ubyte),
        ("colorspace", C.c_ubyte),
        ("qoi_compat", C.c_ubyte),
        ("out_channels", C.c_ubyte),
    ]


class ShardSummary(C.Structure):
    _fields_ = [
        ("first_px", C.c_uint), ("last_px", C.c_uint), ("tail_run", C.c_uint), ("all_run", C.c_uint),
        ("n_px_lo", C.c_uint), ("n_px_hi", C.c_uint), ("slot_valid", C.c_uint * 2), ("slot_px", C.c_uint * 64),
        ("first_slot_px", C.c_uint), ("pad", C.c_uint * 7),
    ]


class DecSummary(C.Structure):
    """``sqoa_b200_dec_summary``."""

    _fields_ = [("exit", C.c_uint), ("has_constant", C.c_uint), ("n_px", C.c_uint), ("val_acc", C.c_uint),
                ("val_flags", C.c_uint), ("needs_serial", C.c_uint), ("pad", C.c_uint * 2)]


class DecCarry(C.Structure):
    """``sqoa_b200_dec_carry``."""

    _fields_ = [("mode", C.c_uint), ("has_carry", C.c_uint), ("entry", C.c_uint), ("pos", C.c_uint),
                ("val_acc", C.c_uint), ("is_last", C.c_uint), ("body_len", C.c_uint), ("n_px", C.c_uint)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class Comm(C.Structure):
    """``sqoa_b200_comm``: rank, world and the caller's all-gather (a callback the library calls in stream order)."""

    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allgather", ALLGATHER_FN), ("user", C.c_void_p)]


DEC_PIXELS, DEC_ENTRY, DEC_SCAN = 0, 1, 2
DEC_SHARD_ALIGN = 1920


class Carry(C.Structure):
    _fields_ = [
        ("has_prev", C.c_uint), ("prev_px", C.c_uint), ("run_in", C.c_uint), ("has_next", C.c_uint),
        ("next_px", C.c_uint), ("slot_px", C.c_uint * 64), ("pad", C.c_uint * 3),
    ]


_lib = None


def lib():
    """The loaded ``libsqoa_b200.so``; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing -- build it with `make -C seqoia_b200/csrc` "
            "(or __graft_entry__.build()); there is no Python or CPU fallback"
        )
    L = C.CDLL(LIB_PATH)
    vp, u, i = C.c_void_p, C.c_uint, C.c_int
    L.sqoa_encode.restype = vp
    L.sqoa_encode.argtypes = [vp, C.POINTER(Desc), C.POINTER(i)]
    L.sqoa_decode.restype = vp
    L.sqoa_decode.argtypes = [vp, i, C.POINTER(Desc), i]
    L.sqoa_write.restype = i
    L.sqoa_write.argtypes = [C.c_char_p, vp, C.POINTER(Desc)]
    L.sqoa_read.restype = vp
    L.sqoa_read.argtypes = [C.c_char_p, C.POINTER(Desc), i]
    L.sqoa_b200_version.restype = C.c_char_p
    L.sqoa_b200_last_error.restype = C.c_char_p
    L.sqoa_b200_max_stream_size.restype = C.c_size_t
    L.sqoa_b200_max_stream_size.argtypes = [u, u, i]
    L.sqoa_b200_probe.restype = i
    L.sqoa_b200_probe.argtypes = [vp, i, C.POINTER(Desc), i, C.POINTER(C.c_longlong)]
    L.sqoa_b200_ctx_create.restype = i
    L.sqoa_b200_ctx_create.argtypes = [C.POINTER(vp), i]
    L.sqoa_b200_ctx_destroy.restype = None
    L.sqoa_b200_ctx_destroy.argtypes = [vp]
    L.sqoa_b200_ctx_set_path.restype = None
    L.sqoa_b200_ctx_set_path.argtypes = [vp, i]
    L.sqoa_b200_ctx_set_qoi_nowait.restype = i
    L.sqoa_b200_ctx_set_qoi_nowait.argtypes = [vp, i]
    L.sqoa_b200_read_many.restype = i
    L.sqoa_b200_read_many.argtypes = [C.POINTER(C.c_char_p), i, i, C.POINTER(vp), C.POINTER(Desc)]
    L.sqoa_b200_write_many.restype = i
    L.sqoa_b200_write_many.argtypes = [C.POINTER(C.c_char_p), i, C.POINTER(vp), C.POINTER(Desc), C.POINTER(i)]
    L.sqoa_b200_host_contexts.restype = i
    L.sqoa_b200_host_contexts.argtypes = []
    L.sqoa_b200_ctx_launch_count.restype = C.c_ulonglong
    L.sqoa_b200_ctx_launch_count.argtypes = [vp]
    L.sqoa_b200_encode_device.restype = i
    L.sqoa_b200_encode_device.argtypes = [vp, vp, C.POINTER(Desc), vp, C.c_size_t, vp, vp]
    L.sqoa_b200_decode_device.restype = i
    L.sqoa_b200_decode_device.argtypes = [vp, vp, i, C.POINTER(Desc), i, vp, C.c_size_t, vp, vp]
    L.sqoa_b200_plan_create.restype = i
    L.sqoa_b200_plan_create.argtypes = [vp, C.POINTER(Item), i, i, C.POINTER(vp)]
    L.sqoa_b200_plan_destroy.restype = None
    L.sqoa_b200_plan_destroy.argtypes = [vp]
    L.sqoa_b200_encode_batch_device.restype = i
    L.sqoa_b200_encode_batch_device.argtypes = [vp, vp, vp, vp, vp, vp]
    L.sqoa_b200_decode_batch_device.restype = i
    L.sqoa_b200_decode_batch_device.argtypes = [vp, vp, vp, vp, vp, vp]
    L.sqoa_b200_shard_summary_device.restype = i
    L.sqoa_b200_shard_summary_device.argtypes = [vp, vp, C.c_ulonglong, i, i, vp, vp]
    L.sqoa_b200_fold_carry.restype = i
    L.sqoa_b200_fold_carry.argtypes = [C.POINTER(ShardSummary), i, i, i, C.POINTER(Carry)]
    L.sqoa_b200_encode_shard_device.restype = i
    L.sqoa_b200_encode_shard_device.argtypes = [vp, vp, C.c_ulonglong, C.POINTER(Desc), vp, vp, C.c_size_t, vp, vp]
    L.sqoa_b200_decode_shard_device.restype = i
    L.sqoa_b200_decode_shard_device.argtypes = [vp, vp, C.c_size_t, C.POINTER(Desc), i, C.POINTER(DecCarry), vp, vp,
                                                C.c_size_t, vp, vp]
    L.sqoa_b200_fold_dec_carry.restype = i
    L.sqoa_b200_fold_dec_carry.argtypes = [C.POINTER(DecSummary), i, i, C.POINTER(DecCarry)]
    L.sqoa_b200_transcode_plan_create.restype = i
    L.sqoa_b200_transcode_plan_create.argtypes = [vp, C.POINTER(Item), i, i, C.POINTER(vp)]
    L.sqoa_b200_transcode_plan_destroy.restype = None
    L.sqoa_b200_transcode_plan_destroy.argtypes = [vp]
    L.sqoa_b200_transcode_batch_device.restype = i
    L.sqoa_b200_transcode_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.sqoa_b200_fold_carry_device.restype = i
    L.sqoa_b200_fold_carry_device.argtypes = [vp, vp, i, i, i, vp, vp]
    L.sqoa_b200_comm_from_nccl.restype = i
    L.sqoa_b200_comm_from_nccl.argtypes = [vp, i, i, C.POINTER(Comm)]
    L.sqoa_b200_encode_sharded_device.restype = i
    L.sqoa_b200_encode_sharded_device.argtypes = [vp, C.POINTER(Comm), vp, C.c_ulonglong, C.POINTER(Desc), vp, C.c_size_t, vp, vp]
    L.sqoa_b200_decode_sharded_device.restype = i
    L.sqoa_b200_decode_sharded_device.argtypes = [vp, C.POINTER(Comm), vp, C.c_size_t, C.c_uint, C.POINTER(Desc), i, vp,
                                                  C.c_size_t, vp, vp, vp]
    _libc = C.CDLL(None)
    _libc.free.argtypes = [vp]
    _libc.free.restype = None
    L._free = _libc.free
    _lib = L
    return L


class SqoaError(RuntimeError):
    pass


def last_error() -> str:
    return lib().sqoa_b200_last_error().decode()


def _check(rc: int, what: str) -> None:
    if rc != OK:
        raise SqoaError(f"{what} failed ({rc}): {last_error()}")


def max_stream_size(width: int, height: int, channels: int) -> int:
    return int(lib().sqoa_b200_max_stream_size(width, height, channels))


def stored_channels(channels: int) -> int:
    return (1 if channels < 3 else 3) + (1 if channels % 2 == 0 else 0)


# ---- the reference's four entry points on host memory --------------------------

def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf).reshape(-1).view(np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def encode(pixels, width: int, height: int, channels: int, colorspace: int = 0, qoi: int = 0) -> Optional[bytes]:
    """``sqoa_encode`` (replaces seqoia.h:363).  ``None`` where the reference returns NULL."""
    L = lib()
    a = _as_u8(pixels)
    d = Desc(width, height, channels, colorspace, qoi)
    n = C.c_int(0)
    p = L.sqoa_encode(a.ctypes.data_as(C.c_void_p), C.byref(d), C.byref(n))
    if not p:
        return None
    out = C.string_at(p, n.value)
    L._free(p)
    return out


def decode(stream, channels: int = 0, size: Optional[int] = None) -> Tuple[Optional[np.ndarray], Desc]:
    """``sqoa_decode`` (replaces seqoia.h:374): (pixels or None, descriptor)."""
    L = lib()
    a = _as_u8(stream)
    n = len(a) if size is None else size
    d = Desc()
    p = L.sqoa_decode(a.ctypes.data_as(C.c_void_p), n, C.byref(d), channels)
    if not p:
        return None, d
    ch = channels if channels != 0 else stored_channels(d.channels)
    out = np.frombuffer(C.string_at(p, d.width * d.height * ch), dtype=np.uint8).copy()
    L._free(p)
    return out, d


def write(filename: str, pixels, width: int, height: int, channels: int, colorspace: int = 0, qoi: int = 0) -> int:
    """``sqoa_write`` (replaces seqoia.h:336): bytes written, 0 on failure."""
    a = _as_u8(pixels)
    d = Desc(width, height, channels, colorspace, qoi)
    return int(lib().sqoa_write(os.fsencode(filename), a.ctypes.data_as(C.c_void_p), C.byref(d)))


def read(filename: str, channels: int = 0) -> Tuple[Optional[np.ndarray], Desc]:
    """``sqoa_read`` (replaces seqoia.h:350)."""
    L = lib()
    d = Desc()
    p = L.sqoa_read(os.fsencode(filename), C.byref(d), channels)
    if not p:
        return None, d
    ch = channels if channels != 0 else stored_channels(d.channels)
    out = np.frombuffer(C.string_at(p, d.width * d.height * ch), dtype=np.uint8).copy()
    L._free(p)
    return out, d


def read_many(filenames, channels: int = 0):
    """``sqoa_b200_read_many``: ``sqoa_read`` for many files in one call (one batch decode per group of files).
    Returns a list of (pixels or None, descriptor)."""
    L = lib()
    n = len(filenames)
    names = (C.c_char_p * n)(*[os.fsencode(f) for f in filenames])
    ptrs = (C.c_void_p * n)()
    descs = (Desc * n)()
    L.sqoa_b200_read_many(names, n, channels, ptrs, descs)
    out = []
    for k in range(n):
        d = Desc(descs[k].width, descs[k].height, descs[k].channels, descs[k].colorspace, descs[k].qoi_compat)
        if not ptrs[k]:
            out.append((None, d))
            continue
        ch = channels if channels != 0 else stored_channels(d.channels)
        out.append((np.frombuffer(C.string_at(ptrs[k], d.width * d.height * ch), dtype=np.uint8).copy(), d))
        L._free(ptrs[k])
    return out


def write_many(filenames, images, descs):
    """``sqoa_b200_write_many``: ``sqoa_write`` for many images in one call.  Returns the bytes written per file (0 on failure)."""
    L = lib()
    n = len(filenames)
    names = (C.c_char_p * n)(*[os.fsencode(f) for f in filenames])
    arrays = [_as_u8(im) for im in images]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    ds = (Desc * n)(*descs)
    sizes = (C.c_int * n)()
    L.sqoa_b200_write_many(names, n, ptrs, ds, sizes)
    return [int(v) for v in sizes]


def host_contexts() -> int:
    """``sqoa_b200_host_contexts``: how many calls of the host entry points run at the same time."""
    return int(lib().sqoa_b200_host_contexts())


def probe(header: bytes, size: int, channels: int = 0) -> Tuple[int, Desc, int]:
    """``sqoa_b200_probe``: (status, descriptor, decoded byte count)."""
    d = Desc()
    n = C.c_longlong(0)
    buf = (C.c_ubyte * 16)(*bytes(header[:15]).ljust(16, b"\0"))
    rc = lib().sqoa_b200_probe(buf, size, C.byref(d), channels, C.byref(n))
    return rc, d, int(n.value)


# ---- device-resident extension surface -------------------------------------------

def _ptr(x) -> int:
    """Device pointer of a torch tensor, or an int passed through."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class Plan:
    def __init__(self, ctx: "Context", items: Sequence[Item], decode_: bool):
        self.ctx = ctx
        self.n = len(items)
        arr = (Item * self.n)(*items)
        h = C.c_void_p()
        _check(lib().sqoa_b200_plan_create(ctx.handle, arr, self.n, 1 if decode_ else 0, C.byref(h)), "plan_create")
        self.handle = h

    def close(self):
        if self.handle:
            lib().sqoa_b200_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TranscodePlan:
    def __init__(self, ctx: "Context", items: Sequence[Item], dst_qoi: int):
        self.ctx = ctx
        self.n = len(items)
        arr = (Item * self.n)(*items)
        h = C.c_void_p()
        _check(lib().sqoa_b200_transcode_plan_create(ctx.handle, arr, self.n, dst_qoi, C.byref(h)), "transcode_plan_create")
        self.handle = h

    def close(self):
        if self.handle:
            lib().sqoa_b200_transcode_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """``sqoa_b200_ctx``: scan workspace + launch bookkeeping for one GPU."""

    def __init__(self, device: int = -1):
        h = C.c_void_p()
        _check(lib().sqoa_b200_ctx_create(C.byref(h), device), "ctx_create")
        self.handle = h

    def close(self):
        if self.handle:
            lib().sqoa_b200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_path(self, path: int) -> None:
        lib().sqoa_b200_ctx_set_path(self.handle, path)

    def set_qoi_nowait(self, on: bool) -> None:
        """``sqoa_b200_ctx_set_qoi_nowait``: QOI decodes queue every stage without reading anything back."""
        _check(lib().sqoa_b200_ctx_set_qoi_nowait(self.handle, 1 if on else 0), "set_qoi_nowait")

    @property
    def launches(self) -> int:
        return int(lib().sqoa_b200_ctx_launch_count(self.handle))

    def encode_device(self, d_pixels, desc: Desc, d_stream, capacity: int, d_len, stream=0) -> None:
        _check(lib().sqoa_b200_encode_device(self.handle, _ptr(d_pixels), C.byref(desc), _ptr(d_stream), capacity,
                                             _ptr(d_len), _ptr(stream)), "encode_device")

    def decode_device(self, d_stream, size: int, desc: Desc, channels: int, d_pixels, capacity: int, d_status=None,
                      stream=0) -> None:
        _check(lib().sqoa_b200_decode_device(self.handle, _ptr(d_stream), size, C.byref(desc), channels,
                                             _ptr(d_pixels), capacity, _ptr(d_status), _ptr(stream)), "decode_device")

    def plan(self, items: Sequence[Item], decode_: bool = False) -> Plan:
        return Plan(self, items, decode_)

    def encode_batch(self, plan: Plan, d_pixels_base, d_streams_base, d_lens, stream=0) -> None:
        _check(lib().sqoa_b200_encode_batch_device(self.handle, plan.handle, _ptr(d_pixels_base),
                                                   _ptr(d_streams_base), _ptr(d_lens), _ptr(stream)), "encode_batch")

    def decode_batch(self, plan: Plan, d_streams_base, d_pixels_base, d_status=None, stream=0) -> None:
        _check(lib().sqoa_b200_decode_batch_device(self.handle, plan.handle, _ptr(d_streams_base),
                                                   _ptr(d_pixels_base), _ptr(d_status), _ptr(stream)), "decode_batch")

    def transcode_plan(self, items: Sequence[Item], dst_qoi: int) -> TranscodePlan:
        """items: decode items (source streams; ``qoi_compat`` = source format) whose ``out_offset`` is where the new stream goes"""
        return TranscodePlan(self, items, dst_qoi)

    def transcode_batch(self, plan: TranscodePlan, d_src_streams, d_dst_streams, d_lens, d_status, stream=0) -> None:
        _check(lib().sqoa_b200_transcode_batch_device(self.handle, plan.handle, _ptr(d_src_streams), _ptr(d_dst_streams),
                                                      _ptr(d_lens), _ptr(d_status), _ptr(stream)), "transcode_batch")

    def fold_carry_device(self, d_summaries, n_shards: int, rank: int, qoi: int, d_carry, stream=0) -> None:
        _check(lib().sqoa_b200_fold_carry_device(self.handle, _ptr(d_summaries), n_shards, rank, qoi, _ptr(d_carry),
                                                 _ptr(stream)), "fold_carry_device")

    def encode_sharded(self, comm: Comm, d_pixels, n_px: int, desc: Desc, d_segment, capacity: int, d_len, stream=0) -> None:
        """``sqoa_b200_encode_sharded_device``: summary, all-gather (``comm.allgather``), device fold, encode -- one call."""
        _check(lib().sqoa_b200_encode_sharded_device(self.handle, C.byref(comm), _ptr(d_pixels), n_px, C.byref(desc),
                                                     _ptr(d_segment), capacity, _ptr(d_len), _ptr(stream)), "encode_sharded")

    def decode_sharded(self, comm: Comm, d_body, avail: int, body_len: int, desc: Desc, channels: int, d_pixels,
                       capacity: int, d_info, d_status, stream=0) -> None:
        """``sqoa_b200_decode_sharded_device``: the three passes of one rank, the all-gathers (``comm.allgather``) and the
        device folds between them -- one call, nothing read back by the host.  ``d_info``: 2 x uint64 on the device
        (first pixel, pixel count); ``d_status``: one int32 on the device."""
        _check(lib().sqoa_b200_decode_sharded_device(self.handle, C.byref(comm), _ptr(d_body), avail, body_len,
                                                     C.byref(desc), channels, _ptr(d_pixels), capacity, _ptr(d_info),
                                                     _ptr(d_status), _ptr(stream)), "decode_sharded")

    def shard_summary(self, d_pixels, n_px: int, channels: int, qoi: int, d_summary, stream=0) -> None:
        _check(lib().sqoa_b200_shard_summary_device(self.handle, _ptr(d_pixels), n_px, channels, qoi,
                                                    _ptr(d_summary), _ptr(stream)), "shard_summary")

    def decode_shard(self, d_body, avail: int, desc: Desc, channels: int, carry: DecCarry, d_summary, d_pixels,
                     capacity: int, d_status=None, stream=0) -> None:
        """One pass (carry.mode) over one byte range of an SQOA stream, see ``sqoa_b200_decode_shard_device``."""
        _check(lib().sqoa_b200_decode_shard_device(self.handle, _ptr(d_body), avail, C.byref(desc), channels,
                                                   C.byref(carry), _ptr(d_summary), _ptr(d_pixels), capacity,
                                                   _ptr(d_status), _ptr(stream)), "decode_shard")

    def encode_shard(self, d_pixels, n_px: int, desc: Desc, d_carry, d_segment, capacity: int, d_len,
                     stream=0) -> None:
        _check(lib().sqoa_b200_encode_shard_device(self.handle, _ptr(d_pixels), n_px, C.byref(desc), _ptr(d_carry),
                                                   _ptr(d_segment), capacity, _ptr(d_len), _ptr(stream)),
               "encode_shard")


def fold_dec_carry(summaries: Sequence[DecSummary], rank: int, carry: DecCarry) -> DecCarry:
    """``sqoa_b200_fold_dec_carry``: fills has_carry / entry / pos / val_acc of ``carry`` for shard ``rank``."""
    arr = (DecSummary * len(summaries))(*summaries)
    _check(lib().sqoa_b200_fold_dec_carry(arr, len(summaries), rank, C.byref(carry)), "fold_dec_carry")
    return carry


def fold_carry(summaries: Sequence[ShardSummary], rank: int, qoi: int) -> Carry:
    arr = (ShardSummary * len(summaries))(*summaries)
    c = Carry()
    _check(lib().sqoa_b200_fold_carry(arr, len(summaries), rank, qoi, C.byref(c)), "fold_carry")
    return c
