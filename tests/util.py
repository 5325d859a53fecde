"""Shared helpers of the test-suite (inputs, golden fixtures, emulator bindings)."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def stored_channels(ch):
    return (1 if ch < 3 else 3) + (1 if ch % 2 == 0 else 0)


def random_image(rng, n, ch, mode):
    """n pixels x ch bytes.  mode 0 noise, 1 smooth walk, 2 long runs, 3 palette (index heavy)."""
    if mode == 0:
        a = rng.integers(0, 256, (n, ch))
    elif mode == 1:
        a = (rng.integers(0, 256, (1, ch)) + np.cumsum(rng.integers(-3, 4, (n, ch)), axis=0)) % 256
    elif mode == 2:
        pal = rng.integers(0, 256, (5, ch))
        reps = rng.integers(1, 1500, n // 7 + 2)
        a = pal[np.resize(np.repeat(rng.integers(0, 5, len(reps)), reps), n)]
    else:
        pal = rng.integers(0, 256, (70, ch))
        a = pal[rng.integers(0, 70, n)]
        a = np.where(rng.random((n, 1)) < 0.3, np.roll(a, 1, axis=0), a)
    return np.ascontiguousarray(a.astype(np.uint8))


def first_difference(a: bytes, b: bytes) -> str:
    n = min(len(a), len(b))
    k = next((i for i in range(n) if a[i] != b[i]), n)
    return f"len {len(a)} vs {len(b)}, first difference at byte {k}: {a[max(0, k - 4):k + 8].hex()} vs {b[max(0, k - 4):k + 8].hex()}"


class Emu:
    """ctypes face of tests/emu/libsqoa_emu.so: the product kernels compiled with -DSQ_EMU."""

    def __init__(self, lib_name="libsqoa_emu.so"):
        self.lib = C.CDLL(os.path.join(ROOT, "tests", "emu", lib_name))
        L = self.lib
        L.emu_configure.argtypes = [C.c_int, C.c_ulonglong]
        L.emu_encode.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.POINTER(C.c_uint)]
        L.emu_encode_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_void_p]
        L.emu_serial.argtypes = [C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_int)]

        L.emu_decode.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.emu_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_void_p]

    def decode_sharded(self, stream, n_px, hdr_channels, out_channels, n_shards, align=1920):
        """SQOA stream decoded as n_shards byte ranges through the three shard passes (ENTRY, SCAN, PIXELS),
        with the host-side fold of tests/util.fold_dec_carry between them; returns the pixels."""
        L = self.lib
        L.emu_decode_shard.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        raw = np.frombuffer(bytes(stream), dtype=np.uint8)
        body = raw[15: len(raw) - 8]
        tiles = (len(body) + align - 1) // align
        per = max(1, (tiles + n_shards - 1) // n_shards)
        cuts = [min(len(body), k * per * align) for k in range(n_shards)] + [len(body)]
        shards = []
        for k in range(n_shards):
            b0, b1 = cuts[k], cuts[k + 1]
            buf = np.zeros(b1 - b0 + 64 + 16, dtype=np.uint8)
            tail = raw[15 + b0: min(len(raw), 15 + b1 + 32)]   # the shard and up to 32 bytes of what follows
            buf[: len(tail)] = tail
            shards.append((buf, b1 - b0, len(tail)))
        def run(k, mode, carry, out=None):
            buf, blen, avail = shards[k]
            c8 = np.array([mode, carry[0], carry[1], carry[2], carry[3], 1 if k == n_shards - 1 else 0, blen, 0],
                          dtype=np.uint32)
            s8 = np.zeros(8, dtype=np.uint32)
            o = out if out is not None else np.zeros(64, dtype=np.uint8)
            st = L.emu_decode_shard(buf.ctypes.data, avail, n_px, hdr_channels, out_channels, c8.ctypes.data,
                                    s8.ctypes.data, o.ctypes.data)
            assert st == 0, (k, mode, st)
            return s8
        summ = [run(k, 1, (0, 0, 0, 0)) for k in range(n_shards)]                     # ENTRY
        carries = [fold_dec_carry(summ, k) for k in range(n_shards)]
        summ = [run(k, 2, (carries[k][0], carries[k][1], 0, 0)) for k in range(n_shards)]  # SCAN with true entries
        carries = [fold_dec_carry(summ, k) for k in range(n_shards)]
        pieces = []
        for k in range(n_shards):
            n_mine = int(summ[k][2]) if k < n_shards - 1 else n_px - carries[k][2]
            out = np.zeros(max(n_mine, 0) * out_channels + 64, dtype=np.uint8)
            run(k, 0, carries[k], out)
            pieces.append(out[: max(n_mine, 0) * out_channels])
        return np.concatenate(pieces)

    def decode_sharded_device_fold(self, stream, n_px, hdr_channels, out_channels, n_shards, align=1920, capacity_px=None):
        """As decode_sharded, but the way sqoa_b200_decode_sharded_device runs it: the carries are folded by the
        device kernel from the gathered summaries and read by the decoder from device memory.  Returns (pixels,
        [(status, first pixel, pixel count)] per shard)."""
        L = self.lib
        L.emu_decode_shard_dev.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_uint, C.c_uint, C.c_uint, C.c_ulonglong, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        raw = np.frombuffer(bytes(stream), dtype=np.uint8)
        body = raw[15: len(raw) - 8]
        tiles = (len(body) + align - 1) // align
        per = max(1, (tiles + n_shards - 1) // n_shards)
        cuts = [min(len(body), k * per * align) for k in range(n_shards)] + [len(body)]
        shards = []
        for k in range(n_shards):
            b0, b1 = cuts[k], cuts[k + 1]
            buf = np.zeros(b1 - b0 + 64 + 16, dtype=np.uint8)
            tail = raw[15 + b0: min(len(raw), 15 + b1 + 32)]
            buf[: len(tail)] = tail
            shards.append((buf, b1 - b0, len(tail)))
        cap = n_px if capacity_px is None else capacity_px
        outs = [np.zeros(cap * out_channels + 64, dtype=np.uint8) for _ in range(n_shards)]
        gathered = np.zeros((n_shards, 8), dtype=np.uint32)
        verdicts = [None] * n_shards
        for mode_next in (1, 2, 0):  # ENTRY, SCAN, PIXELS
            new = np.zeros((n_shards, 8), dtype=np.uint32)
            for k in range(n_shards):
                buf, blen, avail = shards[k]
                info = np.zeros(2, dtype=np.uint64)
                c8 = np.zeros(8, dtype=np.uint32)
                st = L.emu_decode_shard_dev(buf.ctypes.data, avail, n_px, hdr_channels, out_channels, gathered.ctypes.data,
                                            n_shards, k, mode_next, 1 if k == n_shards - 1 else 0, blen, cap,
                                            new[k].ctypes.data, outs[k].ctypes.data, info.ctypes.data, c8.ctypes.data)
                verdicts[k] = (st, int(info[0]), int(info[1]))
            gathered = new
        pieces = [outs[k][: verdicts[k][2] * out_channels] if verdicts[k][0] == 0 else outs[k][:0] for k in range(n_shards)]
        return np.concatenate(pieces), verdicts

    def qoi_decode_sharded(self, stream, n_px, hdr_channels, out_channels, n_shards, capacity_px=None, align=1920):
        """A QOI stream cut into n_shards byte ranges on tile boundaries, decoded range after range the way
        sqoa_b200_decode_sharded_device does on n_shards GPUs.  Returns (pixels put together, [(status, first, count)])."""
        L = self.lib
        L.emu_qoi_decode_sharded.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p]
        raw = np.frombuffer(bytes(stream), dtype=np.uint8).copy()
        body_len = len(raw) - 14 - 8
        tiles = (body_len + align - 1) // align
        per = max(1, (tiles + n_shards - 1) // n_shards)
        inner = [k * per * align for k in range(1, n_shards) if k * per * align < body_len]  # every range but the last is
        cuts = np.array([0] + inner + [body_len], dtype=np.uint32)                           # whole tiles, none is empty
        n_shards = len(cuts) - 1
        cap = n_px if capacity_px is None else capacity_px
        out = np.zeros(n_shards * cap * out_channels + 64, dtype=np.uint8)
        info = np.zeros((n_shards, 2), dtype=np.uint64)
        status = np.full(n_shards, 99, dtype=np.int32)
        rc = L.emu_qoi_decode_sharded(raw.ctypes.data, len(raw), n_px, hdr_channels, out_channels, n_shards, cuts.ctypes.data,
                                      cap, out.ctypes.data, info.ctypes.data, status.ctypes.data)
        assert rc == 0, rc
        verdicts = [(int(status[k]), int(info[k, 0]), int(info[k, 1])) for k in range(n_shards)]
        pieces = [out[k * cap * out_channels: (k * cap + min(verdicts[k][2], cap)) * out_channels] for k in range(n_shards)
                  if verdicts[k][0] == 0]
        return (np.concatenate(pieces) if pieces else out[:0]), verdicts, cuts

    def decode_shard(self, buf, avail, n_px_image, hdr_channels, out_channels, carry, out=None):
        """one pass over one shard; carry: seqoia_b200.DecCarry; returns the 8 summary words"""
        L = self.lib
        L.emu_decode_shard.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        c8 = np.frombuffer(bytes(carry), dtype=np.uint32).copy()
        s8 = np.zeros(8, dtype=np.uint32)
        o = out if out is not None else np.zeros(64, dtype=np.uint8)
        st = L.emu_decode_shard(buf.ctypes.data, avail, n_px_image, hdr_channels, out_channels, c8.ctypes.data,
                                s8.ctypes.data, o.ctypes.data)
        assert st == 0, st
        return s8

    def launch_count(self):
        self.lib.emu_launch_count.restype = C.c_ulonglong
        return int(self.lib.emu_launch_count())

    def qoi_stage_counts(self):
        """QOI decodes that went past the first rows attempt so far: (general pipeline, chained rows attempt, interpreter)"""
        a = (C.c_ulonglong * 3)()
        self.lib.emu_qoi_stage_counts(a)
        return tuple(int(x) for x in a)

    def configure_qoi_fallback(self, whole_group):
        """1: a batch with flagged images is decoded again as a whole; 0: only the flagged images (default)"""
        self.lib.emu_configure_qoi_fallback(int(whole_group))

    def configure_qoi_nowait(self, on):
        """1: QOI decodes queue every stage without reading anything back (sqoa_b200_ctx_set_qoi_nowait); 0: default"""
        self.lib.emu_configure_qoi_nowait(int(on))

    def lanes_stats(self):
        """(tiles the lane-per-chunk QOI tile decoded, tiles it handed to the rows tile) so far"""
        a = (C.c_ulonglong * 2)()
        self.lib.emu_lanes_stats(a)
        return int(a[0]), int(a[1])

    def configure_qoi_lanes(self, off):
        """1: QOI streams without alpha take the rows tile instead of the lane-per-chunk tile; 0: default"""
        self.lib.emu_configure_qoi_lanes(int(off))

    def configure_qoi_rows(self, off):
        """1: QOI decodes skip the one-launch rows kernel (general pipeline only); 0: default"""
        self.lib.emu_configure_qoi_rows(int(off))

    def decode(self, stream, n_px, hdr_channels, qoi, out_channels):
        """parallel decoder; returns (pixels, verdict)"""
        s = np.zeros(len(stream) + 64, dtype=np.uint8)
        s[: len(stream)] = np.frombuffer(bytes(stream), dtype=np.uint8)
        out = np.zeros(n_px * out_channels + 64, dtype=np.uint8)
        st = self.lib.emu_decode(s.ctypes.data, len(stream), n_px, hdr_channels, qoi, out_channels, out.ctypes.data)
        return out[: n_px * out_channels].copy(), st

    def decode_batch(self, streams, n_px, hdr_channels, qoi, out_channels):
        n = len(streams)
        offs = np.zeros(n, dtype=np.uint64)
        sizes = np.array([len(x) for x in streams], dtype=np.uint32)
        pos = 0
        for i, x in enumerate(streams):
            offs[i] = pos
            pos += (len(x) + 15) // 16 * 16 + 3  # deliberately odd alignment
        blob = np.zeros(pos + 64, dtype=np.uint8)
        for i, x in enumerate(streams):
            blob[int(offs[i]): int(offs[i]) + len(x)] = np.frombuffer(bytes(x), dtype=np.uint8)
        stride = n_px * out_channels + 5
        out = np.zeros(n * stride + 64, dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.emu_decode_batch(blob.ctypes.data, offs.ctypes.data, sizes.ctypes.data, n, n_px, hdr_channels, qoi,
                                       out_channels, out.ctypes.data, stride, status.ctypes.data)
        assert rc == 0
        return [out[i * stride: i * stride + n_px * out_channels].copy() for i in range(n)], status

        L.emu_shard_summary.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]

    def shard_summary(self, px, n_px, ch, qoi):
        import seqoia_b200 as sb

        px = np.ascontiguousarray(px, dtype=np.uint8).reshape(-1)
        out = sb.ShardSummary()
        self.lib.emu_shard_summary(px.ctypes.data, n_px, ch, qoi, C.byref(out))
        return out

    def fold_carry_device(self, summaries, rank, qoi):
        """the device fold kernel (what sqoa_b200_fold_carry_device launches)"""
        import seqoia_b200 as sb

        arr = (sb.ShardSummary * len(summaries))(*summaries)
        out = sb.Carry()
        self.lib.emu_fold_carry(C.byref(arr), C.c_int(len(summaries)), C.c_int(rank), C.c_int(qoi), C.byref(out))
        return out

    def configure(self, resident=3, seed=0):
        self.lib.emu_configure(resident, seed)

    def encode(self, img, w, h, ch, qoi, cs=0, flags=3, carry=None, n_px=None):
        img = np.ascontiguousarray(img, dtype=np.uint8).reshape(-1)
        n_px = w * h if n_px is None else n_px
        out = np.zeros(n_px * (ch + 1) + 64, dtype=np.uint8)
        n = C.c_uint(0)
        cptr = None if carry is None else C.cast(C.pointer(carry), C.c_void_p)
        rc = self.lib.emu_encode(img.ctypes.data, n_px, w, h, ch, cs, qoi, flags, cptr, out.ctypes.data, C.byref(n))
        assert rc == 0
        return out[: n.value].tobytes()

    def encode_pieces(self, img, w, h, ch, qoi, piece_tiles):
        img = np.ascontiguousarray(img, dtype=np.uint8).reshape(-1)
        out = np.zeros(w * h * (ch + 1) + 64, dtype=np.uint8)
        n = C.c_uint(0)
        rc = self.lib.emu_encode_pieces(C.c_void_p(img.ctypes.data), C.c_uint(w * h), C.c_uint(w), C.c_uint(h), C.c_int(ch),
                                        C.c_int(qoi), C.c_uint(piece_tiles), C.c_void_p(out.ctypes.data), C.byref(n))
        assert rc == 0
        return out[: n.value].tobytes()

    def decode_pieces(self, stream, n_px, hdr_channels, qoi, out_channels, piece_tiles):
        s = np.zeros(len(stream) + 64, dtype=np.uint8)
        s[: len(stream)] = np.frombuffer(bytes(stream), dtype=np.uint8)
        out = np.zeros(n_px * out_channels + 64, dtype=np.uint8)
        prog = np.zeros(4096, dtype=np.uint32)
        st = self.lib.emu_decode_pieces(C.c_void_p(s.ctypes.data), C.c_uint(len(stream)), C.c_uint(n_px), C.c_int(hdr_channels),
                                        C.c_int(qoi), C.c_int(out_channels), C.c_uint(piece_tiles), C.c_void_p(out.ctypes.data),
                                        C.c_void_p(prog.ctypes.data))
        return out[: n_px * out_channels].copy(), st, prog

    def encode_batch(self, imgs, w, h, ch, qoi):
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
        n = imgs.shape[0]
        stride = w * h * (ch + 1) + 64
        out = np.zeros((n, stride), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.uint32)
        rc = self.lib.emu_encode_batch(imgs.ctypes.data, w * h * ch, n, w, h, ch, qoi, out.ctypes.data, stride,
                                       lens.ctypes.data)
        assert rc == 0
        return [out[i, : lens[i]].tobytes() for i in range(n)]

    def serial_encode(self, img, w, h, ch, qoi, cs=0):
        img = np.ascontiguousarray(img, dtype=np.uint8).reshape(-1)
        out = np.zeros(w * h * (stored_channels(ch) + 1) + 64, dtype=np.uint8)
        n = C.c_uint(0)
        self.lib.emu_serial(0, img.ctypes.data, 0, w, h, ch, cs, qoi, 0, out.ctypes.data, C.byref(n), None)
        return out[: n.value].tobytes()

    def serial_decode(self, stream, w, h, hdr_channels, qoi, out_channels):
        s = np.zeros(len(stream) + 64, dtype=np.uint8)
        s[: len(stream)] = np.frombuffer(bytes(stream), dtype=np.uint8)
        out = np.zeros(w * h * out_channels + 64, dtype=np.uint8)
        st = C.c_int(0)
        self.lib.emu_serial(1, s.ctypes.data, len(stream), w, h, hdr_channels, 0, qoi, out_channels, out.ctypes.data,
                            None, C.byref(st))
        return (None if st.value != 0 else out[: w * h * out_channels].copy()), st.value


def fold_dec_carry(summaries, rank):
    """(has_carry, entry, pos, val_acc) of shard `rank` from the summaries of the shards before it -- the Python
    twin of sqoa_b200_fold_dec_carry (the product's C version is exercised on the GPU)."""
    def badd4(a, b):
        return (((a & 0x7f7f7f7f) + (b & 0x7f7f7f7f)) ^ ((a ^ b) & 0x80808080)) & 0xffffffff
    pos, acc = 0, 0xff000000
    for k in range(rank):
        s = summaries[k]
        assert int(s[5]) == 0, "REF ops"
        assert k == 0 or int(s[1]) == 1, "entry of a shard unknown"
        pos = min(pos + int(s[2]), 0x7fffffff)
        keep = (0x00ffffff if int(s[4]) & 1 else 0) | (0xff000000 if int(s[4]) & 2 else 0)
        acc = (int(s[3]) & keep) | (badd4(acc, int(s[3])) & ~keep & 0xffffffff)
    return (1 if rank > 0 else 0, int(summaries[rank - 1][0]) if rank > 0 else 0, pos, acc)
