"""The N > 1 paths on the CPU: two gloo ranks.  Kernels run in the emulator (this container has no
GPU); the exchange and the host-side fold are the product's own code (seqoia_b200.dist,
sqoa_b200_fold_carry)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, qoi, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    import seqoia_b200 as sb
    from seqoia_b200 import dist as sdist
    from seqoia_b200 import synth
    from util import Emu

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu = Emu()
        # (1) one image sharded by scanline range: only summaries and lengths are exchanged
        w, h, ch = 300, 101, 4
        y0, y1 = sdist.shard_rows(h, world, rank)
        mine = synth.rows("mixed", w, h, ch, y0, y1, seed=11, cell=(40, 9))
        n_px = (y1 - y0) * w
        local = emu.shard_summary(mine, n_px, ch, qoi)
        summaries = sdist.gather_summaries(local)
        carry = sb.fold_carry(summaries, rank, qoi)
        seg = emu.encode(mine, w, h, ch, qoi, 0, flags=4, carry=carry, n_px=n_px)
        lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(lens, torch.tensor([len(seg)], dtype=torch.int64))
        buf = torch.zeros(int(max(x.item() for x in lens)), dtype=torch.uint8)
        buf[: len(seg)] = torch.from_numpy(np.frombuffer(seg, dtype=np.uint8).copy())
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        ok_shard = True
        if rank == 0:
            whole = synth.image("mixed", w, h, ch, seed=11, cell=(40, 9))
            want = oracle.best().encode(whole, w, h, ch, 0, qoi)
            got = b"".join(bytes(parts[r][: int(lens[r].item())].numpy()) for r in range(world))
            ok_shard = got == want
        # (2) a batch sharded by image index: no collective on the data path
        n_img = 7
        lo, hi = sdist.shard_range(n_img, world, rank)
        icons = synth.cfg3(n_img)
        got = emu.encode_batch(icons[lo:hi], 64, 64, 4, qoi)
        ok_batch = all(got[i - lo] == oracle.best().encode(icons[i], 64, 64, 4, 0, qoi) for i in range(lo, hi))
        # (3) one SQOA stream sharded by byte range for decoding: only the 8-word summaries are exchanged
        ok_dec = True
        if qoi == 0:
            whole = synth.image("mixed", w, h, ch, seed=11, cell=(40, 9))
            stream = np.frombuffer(oracle.best().encode(whole, w, h, ch, 0, 0), dtype=np.uint8)
            body_len = len(stream) - 15 - 8
            cuts = sdist.stream_cuts(body_len, world)
            b0, b1 = cuts[rank], cuts[rank + 1]
            tail = stream[15 + b0: min(len(stream), 15 + b1 + 32)]
            buf = np.zeros(len(tail) + 80, dtype=np.uint8)
            buf[: len(tail)] = tail
            carry = sb.DecCarry(sb.DEC_ENTRY, 0, 0, 0, 0, 1 if rank == world - 1 else 0, b1 - b0, 0)

            def gather(words):
                parts = [torch.zeros(8, dtype=torch.int32) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(words.view(np.int32).copy()))
                return [sb.DecSummary.from_buffer_copy(x.numpy().tobytes()) for x in parts]

            summaries = gather(emu.decode_shard(buf, len(tail), w * h, ch, ch, carry))
            sb.fold_dec_carry(summaries, rank, carry)
            carry.mode = sb.DEC_SCAN
            summaries = gather(emu.decode_shard(buf, len(tail), w * h, ch, ch, carry))
            sb.fold_dec_carry(summaries, rank, carry)
            carry.mode = sb.DEC_PIXELS
            n_mine = summaries[rank].n_px if rank < world - 1 else w * h - carry.pos
            out = np.zeros(n_mine * ch + 64, dtype=np.uint8)
            emu.decode_shard(buf, len(tail), w * h, ch, ch, carry, out)
            ok_dec = bool(np.array_equal(out[: n_mine * ch], whole.reshape(-1)[carry.pos * ch: (carry.pos + n_mine) * ch]))
        flags = torch.tensor([int(ok_shard), int(ok_batch), hi - lo, int(ok_dec)], dtype=torch.int64)
        dist.all_reduce(flags, op=dist.ReduceOp.SUM)
        if rank == 0:
            ret["shard"] = ok_shard
            ret["batch_ok"] = int(flags[1].item()) == world
            ret["batch_items"] = int(flags[2].item())
            ret["dec_ok"] = int(flags[3].item()) == world
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("qoi", [0, 1])
def test_two_ranks_gloo_sharded_image_and_batch(qoi):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        port = 29600 + qoi + (os.getpid() % 200) * 2
        procs = [ctx.Process(target=_worker, args=(r, 2, port, qoi, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
            assert p.exitcode == 0
        assert ret["shard"] is True
        assert ret["batch_ok"] is True and ret["batch_items"] == 7
        assert ret["dec_ok"] is True


def test_deal_by_weight_balances_a_mixed_corpus():
    """cfg5 over 8 ranks: every image owned once, pixel counts within a few percent of each other, every rank gets icons
    as well as large images (contiguous index ranges gave one rank all the icons)."""
    from seqoia_b200 import synth
    from seqoia_b200.dist import deal_by_weight

    shapes = synth.cfg5_shapes(1.0)
    px = [w * h for _k, w, h, _c, _s in shapes]
    for world in (2, 4, 8):
        owner = deal_by_weight(px, world)
        assert len(owner) == len(px) and set(owner) == set(range(world))
        load = [sum(p for p, o in zip(px, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(px), (world, load)
        small = [sum(1 for p, o in zip(px, owner) if o == r and p <= 64 * 64) for r in range(world)]
        assert min(small) > 0, small
    assert deal_by_weight([5, 5, 5], 1) == [0, 0, 0]


def test_shard_ranges_cover_everything_once():
    from seqoia_b200.dist import shard_range

    for n in (1, 7, 8, 100_000, 19999):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
