"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the exchange.

* Image batches shard by image index with **no collective** (:func:`shard_range`).
* One large image shards by scanline range; the only data that crosses GPUs is the boundary
  summary of every shard (320 bytes: first / last pixel, open run, the 64 QOI index slots),
  all-gathered once (NCCL on GPUs, gloo in the CPU tests), plus the segment lengths.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np

from . import (DEC_ENTRY, DEC_PIXELS, DEC_SCAN, DEC_SHARD_ALIGN, Carry, DecCarry, DecSummary, Desc, ShardSummary,
               fold_carry, fold_dec_carry)


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` of ``n_items`` for ``rank`` (no communication needed)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def deal_by_weight(weights, world: int) -> List[int]:
    """Owner rank of every item when independent items of very different cost (a mixed corpus: icons next to 8 Mpx
    screenshots) are spread over ``world`` ranks without communication: largest first, each to the rank with the least
    weight so far.  Deterministic, so every rank computes the same assignment on its own."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0] * world
    owner = [0] * len(weights)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += weights[i]
    return owner


def shard_rows(height: int, world: int, rank: int) -> Tuple[int, int]:
    """Scanline range of one image owned by ``rank``."""
    return shard_range(height, world, rank)


def summary_to_array(s: ShardSummary) -> np.ndarray:
    return np.frombuffer(bytes(s), dtype=np.int32).copy()


def array_to_summary(a: np.ndarray) -> ShardSummary:
    return ShardSummary.from_buffer_copy(np.ascontiguousarray(a, dtype=np.int32).tobytes())


def gather_summaries(local, group=None) -> List[ShardSummary]:
    """All-gather the per-shard summaries.  ``local`` is a torch int32 tensor of 80 words (on the
    GPU for NCCL, on the CPU for gloo) or a :class:`ShardSummary`."""
    import torch
    import torch.distributed as dist

    if isinstance(local, ShardSummary):
        local = torch.from_numpy(summary_to_array(local))
    world = dist.get_world_size(group)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local, group=group)
    return [array_to_summary(t.cpu().numpy()) for t in out]


def torch_comm(group=None) -> "Comm":
    """A :class:`Comm` whose all-gather is ``torch.distributed.all_gather_into_tensor`` on the caller's stream (NCCL for
    CUDA tensors): what ``sqoa_b200_encode_sharded_device`` calls between the summary kernels and the device fold."""
    import torch
    import torch.distributed as dist

    from . import ALLGATHER_FN, Comm

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0

    # First choice: hand the library torch's own NCCL communicator, so that the all-gather is one ncclAllGather call
    # from C on the caller's stream (sqoa_b200_comm_from_nccl) with no Python in the data path.
    if world > 1 and os.environ.get("SQOA_B200_PY_ALLGATHER") != "1":
        try:
            from . import lib

            pg = (group or dist.distributed_c10d._get_default_group())._get_backend(torch.device("cuda"))
            ptr = int(pg._comm_ptr())
            comm = Comm()
            if ptr and lib().sqoa_b200_comm_from_nccl(ptr, rank, world, C.byref(comm)) == 0:
                comm._native = True
                return comm
        except Exception:
            pass

    cache = {}  # (pointer, bytes) -> tensor view; stream handle -> torch stream (the library's buffers do not move)

    def allgather(_user, d_send, d_recv, nbytes, cuda_stream):
        try:
            key = (d_send, d_recv, nbytes)
            if key not in cache:
                dev = torch.device("cuda", torch.cuda.current_device())
                cache[key] = (_wrap_device_bytes(d_send, nbytes, dev), _wrap_device_bytes(d_recv, nbytes * world, dev))
            send, recv = cache[key]
            skey = ("stream", cuda_stream or 0)
            if skey not in cache:
                cache[skey] = torch.cuda.ExternalStream(cuda_stream or 0)
            with torch.cuda.stream(cache[skey]):
                dist.all_gather_into_tensor(recv, send, group=group)
            return 0
        except Exception:  # reported to the C caller as a failed collective
            import traceback

            traceback.print_exc()
            return 1

    cb = ALLGATHER_FN(allgather)
    comm = Comm(rank, world, cb, None)
    comm._keepalive = cb
    return comm


def _wrap_device_bytes(ptr: int, nbytes: int, dev):
    """A uint8 torch view of device memory owned by someone else (the library's exchange buffers)."""
    import torch

    class _Mem:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_Mem(), device=dev)


class ShardedEncoder:
    """Scanline-sharded encode of one image, one library call per rank: boundary summary, all-gather of the 320-byte
    summaries, fold on the device, encode -- all stream-ordered, no host round trip."""

    def __init__(self, ctx, device, group=None):
        self.ctx = ctx
        self.comm = torch_comm(group)

    def encode(self, d_pixels, n_px: int, desc: Desc, d_segment, capacity: int, d_len, stream=0) -> None:
        self.ctx.encode_sharded(self.comm, d_pixels, n_px, desc, d_segment, capacity, d_len, stream)


def encode_sharded_device(ctx, d_pixels, n_px: int, desc: Desc, d_segment, capacity: int, d_len, group=None,
                          stream=0):
    """Encode this rank's scanline shard of one image (device resident), the step-by-step way (summary, all-gather,
    HOST fold, encode); :class:`ShardedEncoder` is the one-call, device-fold form.

    Returns ``(carry, summaries)``; the segment is in ``d_segment[:d_len]``.  Rank 0's segment
    starts with the header, the last rank's ends with the end marker; concatenated in rank order
    the segments are the reference's stream."""
    import torch
    import torch.distributed as dist

    stored = 3 if desc.channels in (3, 5) else 4
    with torch.cuda.stream(torch.cuda.ExternalStream(stream or 0)) if stream else _nullcontext():
        d_sum = torch.zeros(80, dtype=torch.int32, device=d_pixels.device)
        ctx.shard_summary(d_pixels, n_px, stored, desc.qoi_compat, d_sum, stream)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            summaries = gather_summaries(d_sum, group)
            rank = dist.get_rank(group)
        else:
            torch.cuda.synchronize()
            summaries = [array_to_summary(d_sum.cpu().numpy())]
            rank = 0
        carry = fold_carry(summaries, rank, desc.qoi_compat)
        d_carry = torch.from_numpy(np.frombuffer(bytes(carry), dtype=np.int32).copy()).to(d_pixels.device)
        ctx.encode_shard(d_pixels, n_px, desc, d_carry, d_segment, capacity, d_len, stream)
    return carry, summaries, d_carry


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


# ---- stream-sharded decode of one SQOA image -------------------------------------------------------
def stream_cuts(body_len: int, world: int) -> List[int]:
    """Byte offsets (in the op stream, i.e. after the 15-byte header) at which the shards start: equal numbers
    of decoder tiles, every cut on a tile boundary; ``cuts[world] == body_len``."""
    tiles = (body_len + DEC_SHARD_ALIGN - 1) // DEC_SHARD_ALIGN
    per, extra = divmod(tiles, world)
    cuts, t = [], 0
    for r in range(world):
        cuts.append(min(body_len, t * DEC_SHARD_ALIGN))
        t += per + (1 if r < extra else 0)
    cuts.append(body_len)
    return cuts


def gather_dec_summaries(local: "torch.Tensor", group=None) -> List[DecSummary]:
    """All-gather of the 8-word shard summaries (``local``: int32 tensor of 8 words on the rank's device)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        parts = [local]
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
    return [DecSummary.from_buffer_copy(p.cpu().numpy().astype(np.int32).tobytes()) for p in parts]


def decode_stream_shard(ctx, d_body, avail: int, body_len: int, desc: Desc, channels: int, rank: int, world: int,
                        d_summary, alloc_pixels, stream=0, group=None):
    """The three passes of one rank: ENTRY -> all-gather -> SCAN -> all-gather -> PIXELS.  ``d_summary`` is an int32
    tensor of 8 words on the device; ``alloc_pixels(n_bytes)`` returns the device buffer for this shard's pixels.
    Returns (pixel buffer, first pixel index, number of pixels)."""
    import torch

    carry = DecCarry(DEC_ENTRY, 0, 0, 0, 0, 1 if rank == world - 1 else 0, body_len, 0)
    ctx.decode_shard(d_body, avail, desc, channels, carry, d_summary, None, 0, None, stream)
    summaries = gather_dec_summaries(d_summary, group)
    fold_dec_carry(summaries, rank, carry)
    carry.mode = DEC_SCAN
    ctx.decode_shard(d_body, avail, desc, channels, carry, d_summary, None, 0, None, stream)
    summaries = gather_dec_summaries(d_summary, group)
    fold_dec_carry(summaries, rank, carry)
    carry.mode = DEC_PIXELS
    n_total = desc.width * desc.height
    oc = channels if channels else (4 if desc.channels % 2 == 0 else 3)
    n_mine = summaries[rank].n_px if rank < world - 1 else max(0, n_total - carry.pos)
    n_mine = min(n_mine, max(0, n_total - carry.pos))
    d_px = alloc_pixels(n_mine * oc + 64)
    ctx.decode_shard(d_body, avail, desc, channels, carry, None, d_px, n_mine * oc + 64, None, stream)
    return d_px, carry.pos, n_mine


class ShardedDecoder:
    """Stream-sharded decode of one image, one library call per rank (``sqoa_b200_decode_sharded_device``).  SQOA: the
    three passes, the two all-gathers of 32-byte summaries and the device folds between them are all stream-ordered;
    the host reads nothing back in between.  QOI: the ranges are decoded one after the other, each from the 544-byte
    carry of the one before it (``d_body`` 16-byte aligned, 64 bytes of look-ahead; see the header).  The pixel buffer is kept between calls; its size is a guess the first
    time (the shard's pixel count is only known on the device) and grows when the library reports that it was too
    small.  ``step_by_step=True`` keeps the older form with host folds (:func:`decode_stream_shard`)."""

    def __init__(self, ctx, device, group=None, step_by_step: bool = False):
        import torch

        self.ctx, self.group = ctx, group
        self.d_sum = torch.zeros(8, dtype=torch.int32, device=device)
        self.d_info = torch.zeros(2, dtype=torch.int64, device=device)
        self.d_status = torch.zeros(1, dtype=torch.int32, device=device)
        self.buf = None
        self.device = device
        self.step_by_step = step_by_step
        self.comm = None if step_by_step else torch_comm(group)

    def _alloc(self, nbytes):
        import torch

        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.buf

    def launch(self, d_body, avail: int, body_len: int, desc: Desc, channels: int, stream=0):
        """Asynchronous: queues the whole decode on ``stream``; pixels land in ``self.buf``, the verdict in
        ``self.d_status`` and (first pixel, count) in ``self.d_info``."""
        oc = channels if channels else (4 if desc.channels % 2 == 0 else 3)
        if self.buf is None:
            world = max(1, self.comm.world)
            self._alloc((desc.width * desc.height * oc) // world * 2 + 4096)
        self.ctx.decode_sharded(self.comm, d_body, avail, body_len, desc, channels, self.buf, self.buf.numel(), self.d_info,
                                self.d_status, stream)

    def decode(self, d_body, avail: int, body_len: int, desc: Desc, channels: int, rank: int, world: int, stream=0):
        """Returns (pixel buffer, first pixel index, number of pixels) of this rank's shard."""
        import torch

        from . import SqoaError

        if self.step_by_step:
            return decode_stream_shard(self.ctx, d_body, avail, body_len, desc, channels, rank, world, self.d_sum,
                                       self._alloc, stream, self.group)
        oc = channels if channels else (4 if desc.channels % 2 == 0 else 3)
        for _attempt in range(2):
            self.launch(d_body, avail, body_len, desc, channels, stream)
            if stream:
                torch.cuda.ExternalStream(stream).synchronize()
            else:
                torch.cuda.current_stream().synchronize()
            st = int(self.d_status.item())
            first, count = (int(v) for v in self.d_info.tolist())
            again = 1 if st == -3 else 0
            if self.comm.world > 1:  # the call is collective: every rank repeats it if one has to
                import torch.distributed as dist

                flag = torch.tensor([again], dtype=torch.int32, device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
                again = int(flag.item())
            if st == -3:  # the guess was too small: the library says what the shard needs
                self.buf = None
                self._alloc(count * oc + 4096)
            if again:
                continue
            if st:
                raise SqoaError(f"decode_sharded: stream rejected ({st})")
            return self.buf, first, count
        raise SqoaError("decode_sharded: pixel buffer still too small")

    def describe(self, qoi: bool = False) -> str:
        if qoi:
            return ("byte ranges decoded one after the other, the decoder's state (64 slots, running pixel, entry, hash, "
                    "pixel count: 544 bytes) all-gathered after every range")
        if self.step_by_step:
            return "entry + scan + pixels, two all-gathers of 32-byte summaries, host folds"
        return "entry + scan + pixels in one call, two all-gathers of 32-byte summaries folded on the device"
