"""Parity on the GPU, through the C ABI (ctypes -> libsqoa_b200.so): every stream and
every decoded pixel buffer is byte-compared with the oracle (the compiled reference
when oracle/_ref travelled with the snapshot, else the pinned restatement) and with
the committed reference digests."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
import seqoia_b200 as sb
from seqoia_b200 import synth
from util import first_difference, golden, random_image, stored_channels


@pytest.fixture(scope="module")
def cpu():
    return oracle.best()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a B200; there is no CPU fallback to test"
    torch.cuda.set_device(0)
    return torch


def test_golden_encode_vectors(torch_cuda):
    for v in golden("kat.json")["encode"]:
        px = np.frombuffer(bytes.fromhex(v["pixels"]), dtype=np.uint8)
        got = sb.encode(px, v["w"], v["h"], v["channels"], v["colorspace"], v["qoi"])
        want = None if v["stream"] is None else bytes.fromhex(v["stream"])
        assert got == want, v["name"]


def test_golden_decode_vectors(torch_cuda):
    for v in golden("kat.json")["decode"]:
        px, d = sb.decode(bytes.fromhex(v["stream"]), v["channels"])
        want = None if v["pixels"] is None else bytes.fromhex(v["pixels"])
        assert (None if px is None else px.tobytes()) == want, v["name"]
        if want is not None:
            assert [d.width, d.height, d.channels, d.colorspace, d.qoi_compat] == v["desc"], v["name"]


@pytest.mark.parametrize("qoi", [0, 1])
def test_random_images_all_channel_layouts(torch_cuda, cpu, qoi):
    rng = np.random.default_rng(40 + qoi)
    for it in range(150):
        ch = int(rng.integers(1, 7))
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 60))
        if it % 10 == 0:
            w, h = 4099, 3
        img = random_image(rng, w * h, stored_channels(ch), it % 4)
        want = cpu.encode(img, w, h, ch, it & 1, qoi)
        got = sb.encode(img, w, h, ch, it & 1, qoi)
        if want is None:
            assert got is None
            continue
        assert got == want, (it, w, h, ch, first_difference(got, want))
        for oc in (0, 3, 4) if it % 3 else (0, 1, 2, 3, 4):
            px, d = sb.decode(got, oc)
            ref_px, ref_d = cpu.decode(got, oc)
            assert np.array_equal(px, ref_px), (it, oc)
            assert (d.width, d.height, d.channels, d.colorspace, d.qoi_compat) == (
                ref_d.width, ref_d.height, ref_d.channels, ref_d.colorspace, ref_d.qoi_compat)


@pytest.mark.parametrize("name,maker", [("cfg1_1920x1080_rgba", lambda: synth.cfg1()),
                                        ("cfg2_3840x2160_rgb", lambda: synth.cfg2()),
                                        ("cfg2_3840x2160_rgba", lambda: synth.cfg2(channels=4)),
                                        ("cfg4_scaled_2000x1999", lambda: synth.cfg4(2000, 1999)),
                                        ("screen_1280x720_rgb", lambda: synth.image("screen", 1280, 720, 3, seed=7, cell=(160, 90)))])
@pytest.mark.parametrize("qoi", [0, 1])
def test_full_size_configs_match_reference_digests(torch_cuda, name, maker, qoi):
    """BASELINE.json configs at full size against the digests of the reference's own streams."""
    dig = golden("digests.json")["digests"][f"{name}_q{qoi}"]
    img = maker()
    assert hashlib.sha256(img.tobytes()).hexdigest() == dig["pixels_sha256"], "generator drifted"
    s = sb.encode(img, dig["w"], dig["h"], dig["channels"], 0, qoi)
    assert s is not None and len(s) == dig["stream_len"]
    assert hashlib.sha256(s).hexdigest() == dig["stream_sha256"]


@pytest.mark.parametrize("qoi", [0, 1])
def test_device_entry_points_and_paths_agree(torch_cuda, cpu, qoi):
    torch = torch_cuda
    ctx = sb.Context(0)
    img = synth.image("mixed", 700, 300, 4, seed=9)
    want = cpu.encode(img, 700, 300, 4, 0, qoi)
    cap = sb.max_stream_size(700, 300, 4)
    d_px = torch.from_numpy(img.reshape(-1)).cuda()
    outs = []
    for path in (sb.PATH_PARALLEL, sb.PATH_SERIAL):
        ctx.set_path(path)
        d_s = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
        ctx.encode_device(d_px, sb.Desc(700, 300, 4, 0, qoi), d_s, cap, d_n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append(bytes(d_s[: int(d_n.item())].cpu().numpy()))
    assert outs[0] == want, first_difference(outs[0], want)
    assert outs[1] == want, first_difference(outs[1], want)
    with pytest.raises(sb.SqoaError):
        ctx.encode_device(d_px, sb.Desc(700, 300, 4, 0, qoi), d_s, cap - 1, d_n, 0)  # capacity check


@pytest.mark.parametrize("qoi", [0, 1])
def test_batch_of_icons(torch_cuda, cpu, qoi):
    torch = torch_cuda
    n = 512
    icons = synth.cfg3(n)
    cap = (sb.max_stream_size(64, 64, 4) + 63) // 64 * 64
    items = [sb.Item(i * 64 * 64 * 4, i * cap, 64, 64, 0, 4, 0, qoi, 0) for i in range(n)]
    ctx = sb.Context(0)
    plan = ctx.plan(items)
    d_px = torch.from_numpy(icons.reshape(-1)).cuda()
    d_out = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.encode_batch(plan, d_px, d_out, d_len, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    lens = d_len.cpu().numpy()
    out = d_out.cpu().numpy()
    h = hashlib.sha256()
    for i in range(n):
        got = out[i * cap: i * cap + lens[i]].tobytes()
        if i < 64:
            assert got == cpu.encode(icons[i], 64, 64, 4, 0, qoi), i
        if i < 256:
            h.update(got)
    dig = golden("digests.json")["digests"][f"cfg3_icons_0_255_q{qoi}"]
    assert h.hexdigest() == dig["stream_sha256"]


def test_write_and_read_files(torch_cuda, cpu, tmp_path):
    img = synth.image("icon", 200, 100, 4, seed=5)
    for qoi in (0, 1):
        path = str(tmp_path / f"x{qoi}.sqoa")
        n = sb.write(path, img, 200, 100, 4, 1, qoi)
        data = open(path, "rb").read()
        assert n == len(data) and data == cpu.encode(img, 200, 100, 4, 1, qoi)
        px, d = sb.read(path, 0)
        assert np.array_equal(px, img.reshape(-1)) and d.colorspace == 1 and d.qoi_compat == qoi
        px3, _ = sb.read(path, 3)
        assert np.array_equal(px3, img.reshape(-1, 4)[:, :3].reshape(-1))
    assert sb.write(str(tmp_path / "no" / "dir" / "x"), img, 200, 100, 4) == 0
    assert sb.read(str(tmp_path / "missing"))[0] is None


def test_round_trip_property_on_large_random_image(torch_cuda):
    """Size-independent property: decode(encode(x)) == x, 4 Mpx of mixed content, both formats."""
    img = synth.image("mixed", 2500, 1601, 4, seed=77)
    for qoi in (0, 1):
        s = sb.encode(img, 2500, 1601, 4, 0, qoi)
        px, d = sb.decode(s, 0)
        assert np.array_equal(px, img.reshape(-1))


@pytest.mark.parametrize("qoi", [0, 1])
def test_scanline_shards_on_one_gpu(torch_cuda, cpu, qoi):
    """cfg4's mechanism on one GPU: three shards of one image, summaries -> fold -> shard encode;
    concatenated segments == the reference's stream (the NCCL exchange itself is covered by
    bench.py --workload cfg4 under torchrun and by the gloo test)."""
    torch = torch_cuda
    from seqoia_b200 import dist as sdist

    w, h = 2000, 1999
    img = synth.cfg4(w, h)
    ctx = sb.Context(0)
    s = torch.cuda.current_stream().cuda_stream
    world = 3
    shards = []
    for r in range(world):
        y0, y1 = sdist.shard_rows(h, world, r)
        d_px = torch.from_numpy(img[y0:y1].reshape(-1)).cuda()
        d_sum = torch.zeros(80, dtype=torch.int32, device="cuda")
        ctx.shard_summary(d_px, (y1 - y0) * w, 4, qoi, d_sum, s)
        shards.append((d_px, (y1 - y0) * w, d_sum))
    torch.cuda.synchronize()
    summaries = [sdist.array_to_summary(x[2].cpu().numpy()) for x in shards]
    out = b""
    for r, (d_px, n_px, _) in enumerate(shards):
        carry = sb.fold_carry(summaries, r, qoi)
        d_carry = torch.from_numpy(np.frombuffer(bytes(carry), dtype=np.int32).copy()).cuda()
        cap = n_px * 5 + 64
        d_seg = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
        ctx.encode_shard(d_px, n_px, sb.Desc(w, h, 4, 0, qoi), d_carry, d_seg, cap, d_len, s)
        torch.cuda.synchronize()
        out += bytes(d_seg[: int(d_len.item())].cpu().numpy())
    dig = golden("digests.json")["digests"][f"cfg4_scaled_2000x1999_q{qoi}"]
    assert len(out) == dig["stream_len"] and hashlib.sha256(out).hexdigest() == dig["stream_sha256"]


@pytest.mark.parametrize("qoi", [0, 1])
def test_batch_decode_of_icons(torch_cuda, cpu, qoi):
    torch = torch_cuda
    n = 300
    icons = synth.cfg3(n)
    streams = [cpu.encode(icons[i], 64, 64, 4, 0, qoi) for i in range(n)]
    offs, pos = [], 0
    for x in streams:
        offs.append(pos)
        pos += (len(x) + 63) // 64 * 64
    blob = np.zeros(pos + 64, dtype=np.uint8)
    for o, x in zip(offs, streams):
        blob[o: o + len(x)] = np.frombuffer(x, dtype=np.uint8)
    items = [sb.Item(offs[i], i * 16384, 64, 64, len(streams[i]), 4, 0, qoi, 4) for i in range(n)]
    ctx = sb.Context(0)
    plan = ctx.plan(items, decode_=True)
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(n * 16384, dtype=torch.uint8, device="cuda")
    d_st = torch.ones(n, dtype=torch.int32, device="cuda")
    ctx.decode_batch(plan, d_in, d_out, d_st, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0
    assert np.array_equal(d_out.cpu().numpy(), icons.reshape(-1))


@pytest.mark.gpu
def test_batch_decode_of_mixed_qoi_images(torch_cuda, cpu):
    """Opaque photo-like images (one-launch rows kernel), images whose alpha moves and an image that reads a
    never-written slot (handed on, alone, to the general pipeline) in ONE batch; every pixel as the reference's."""
    torch = torch_cuda
    rng = np.random.default_rng(99)
    w, h = 640, 360
    imgs = []
    for i in range(9):
        img = synth.image("photo", w, h, 4, seed=300 + i).reshape(-1, 4).copy()
        if i % 3 == 1:
            img[rng.random(w * h) < 0.05, 3] = 90          # RGBA ops
        if i == 5:
            img[: w * h // 3] = 0                          # starts with transparent black: INDEX 0 before any write
        imgs.append(img.reshape(-1))
    streams = [cpu.encode(im, w, h, 4, 0, 1) for im in imgs]
    offs, pos = [], 0
    for x in streams:
        offs.append(pos)
        pos += (len(x) + 63) // 64 * 64 + 5
    blob = np.zeros(pos + 64, dtype=np.uint8)
    for o, x in zip(offs, streams):
        blob[o: o + len(x)] = np.frombuffer(x, dtype=np.uint8)
    stride = w * h * 4
    items = [sb.Item(offs[i], i * stride, w, h, len(streams[i]), 4, 0, 1, 4) for i in range(len(streams))]
    ctx = sb.Context(0)
    plan = ctx.plan(items, decode_=True)
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(len(streams) * stride, dtype=torch.uint8, device="cuda")
    d_st = torch.ones(len(streams), dtype=torch.int32, device="cuda")
    for _ in range(2):   # twice: the flag counter and the sub-table are reused
        d_out.zero_()
        ctx.decode_batch(plan, d_in, d_out, d_st, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert int(d_st.abs().sum().item()) == 0
        got = d_out.cpu().numpy().reshape(len(streams), stride)
        for i in range(len(streams)):
            want, _ = cpu.decode(streams[i], 4)
            assert np.array_equal(got[i], want), i


@pytest.mark.gpu
@pytest.mark.parametrize("ch", [3, 4])
def test_stream_sharded_decode_on_one_gpu(torch_cuda, cpu, ch):
    """SURVEY 8e, single image decode: 1..5 byte ranges of one SQOA stream through the three shard passes of the
    C ABI (ENTRY / SCAN / PIXELS) and sqoa_b200_fold_dec_carry; the pieces put together equal the whole decode."""
    torch = torch_cuda
    w, h = 1531, 420
    img = synth.image("mixed", w, h, ch, seed=5, cell=(61, 23))
    stream = np.frombuffer(cpu.encode(img, w, h, ch, 0, 0), dtype=np.uint8)
    want = img.reshape(-1)
    body_len = len(stream) - 15 - 8
    ctx = sb.Context(0)
    desc = sb.Desc(w, h, ch, 0, 0)
    from seqoia_b200 import dist as sdist

    for world in (1, 2, 5):
        cuts = sdist.stream_cuts(body_len, world)
        bufs, carries = [], []
        for r in range(world):
            b0, b1 = cuts[r], cuts[r + 1]
            tail = stream[15 + b0: min(len(stream), 15 + b1 + 32)]
            bufs.append((torch.from_numpy(tail.copy()).cuda(), len(tail)))
            carries.append(sb.DecCarry(sb.DEC_ENTRY, 0, 0, 0, 0, 1 if r == world - 1 else 0, b1 - b0, 0))
        d_sum = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(world)]

        def run_all(mode, pixels=None):
            for r in range(world):
                carries[r].mode = mode
                d_px = pixels[r] if pixels else None
                ctx.decode_shard(bufs[r][0], bufs[r][1], desc, 0, carries[r], d_sum[r] if pixels is None else None, d_px,
                                 0 if d_px is None else d_px.numel(), None, 0)
            torch.cuda.synchronize()
            return [sb.DecSummary.from_buffer_copy(x.cpu().numpy().tobytes()) for x in d_sum]

        for mode in (sb.DEC_ENTRY, sb.DEC_SCAN):
            sums = run_all(mode)
            for r in range(world):
                sb.fold_dec_carry(sums, r, carries[r])
        counts = [sums[r].n_px if r < world - 1 else w * h - carries[r].pos for r in range(world)]
        pixels = [torch.zeros(counts[r] * ch + 64, dtype=torch.uint8, device="cuda") for r in range(world)]
        run_all(sb.DEC_PIXELS, pixels)
        got = np.concatenate([pixels[r][: counts[r] * ch].cpu().numpy() for r in range(world)])
        assert np.array_equal(got, want), world


@pytest.mark.gpu
@pytest.mark.parametrize("nowait", [0, 1])
def test_qoi_batch_whose_images_take_every_decode_attempt(torch_cuda, cpu, nowait):
    """Opaque photos (first attempt of the rows kernel), half-transparent palette images whose alpha guesses fail
    (second attempt, tiles chained) and a hand-made stream that reads never-written slots (general pipeline), in one
    batch and one by one; every pixel as the reference's.  nowait = 1: the same through the mode that reads nothing
    back between the stages (sqoa_b200_ctx_set_qoi_nowait), twice, then the default mode again on the same context."""
    torch = torch_cuda
    rng = np.random.default_rng(123)
    w, h = 640, 200
    n_px = w * h

    def half_transparent():
        pal = rng.integers(0, 256, (40, 3), dtype=np.uint8)
        img = np.zeros((n_px, 4), np.uint8)
        img[:, :3] = pal[rng.integers(0, 40, n_px)]
        fresh = rng.random(n_px) < 0.3
        img[fresh, :3] = rng.integers(0, 256, (int(fresh.sum()), 3))
        img[:, 3] = 128
        return img.reshape(-1)

    streams = [cpu.encode(synth.image("photo", w, h, 4, seed=40 + i).reshape(-1), w, h, 4, 0, 1) for i in range(3)]
    streams += [cpu.encode(half_transparent(), w, h, 4, 0, 1) for _ in range(2)]
    hdr = b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([4, 0])
    body = bytes([0xFE, 10, 20, 30, 0x05, 0xFE, 1, 2, 3, 0x05, 0xC1]) + bytes([0xFD]) * 2500 + bytes([0x07, 0xFE, 5, 6, 7, 0x07])
    streams.append(hdr + body + bytes(7) + b"\x01")
    streams.append(cpu.encode(synth.image("icon", w, h, 4, seed=77).reshape(-1), w, h, 4, 0, 1))
    want = [cpu.decode(s, 4)[0] for s in streams]
    offs, pos = [], 0
    for x in streams:
        offs.append(pos)
        pos += (len(x) + 63) // 64 * 64 + 3
    blob = np.zeros(pos + 64, dtype=np.uint8)
    for o, x in zip(offs, streams):
        blob[o: o + len(x)] = np.frombuffer(x, dtype=np.uint8)
    stride = n_px * 4
    items = [sb.Item(offs[i], i * stride, w, h, len(streams[i]), 4, 0, 1, 4) for i in range(len(streams))]
    ctx = sb.Context(0)
    plan = ctx.plan(items, decode_=True)
    d_in = torch.from_numpy(blob).cuda()
    for mode in ([0] if not nowait else [1, 1, 0]):
        ctx.set_qoi_nowait(bool(mode))
        d_out = torch.zeros(len(streams) * stride, dtype=torch.uint8, device="cuda")
        d_st = torch.ones(len(streams), dtype=torch.int32, device="cuda")
        before = ctx.launches
        ctx.decode_batch(plan, d_in, d_out, d_st, torch.cuda.current_stream().cuda_stream)
        if mode:
            assert ctx.launches - before == 6
        torch.cuda.synchronize()
        assert int(d_st.abs().sum().item()) == 0, mode
        got = d_out.cpu().numpy().reshape(len(streams), stride)
        for i in range(len(streams)):
            assert np.array_equal(got[i], want[i]), (mode, i)
        if mode:  # single images through the device entry point in the same mode
            for i in (0, 3, 5):
                d_one = torch.zeros(stride, dtype=torch.uint8, device="cuda")
                d_s1 = torch.ones(1, dtype=torch.int32, device="cuda")
                rc, dd, nb = sb.probe(streams[i][:15], len(streams[i]), 4)
                ctx.decode_device(torch.from_numpy(np.frombuffer(streams[i], dtype=np.uint8).copy()).cuda(), len(streams[i]), dd,
                                  4, d_one, stride, d_s1, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert int(d_s1.item()) == 0 and np.array_equal(d_one.cpu().numpy(), want[i]), (mode, i)
    for i, s in enumerate(streams):   # and through the host entry point, one by one
        px, _d = sb.decode(s, 4)
        assert px is not None and np.array_equal(px, want[i]), i


@pytest.mark.gpu
def test_many_files_in_one_call(torch_cuda, cpu, tmp_path):
    """sqoa_b200_write_many / sqoa_b200_read_many (the file path with many-file batching): a mix of sizes,
    channel layouts (mono included) and both formats written in one call == the reference's streams byte for byte; read
    back in one call == the reference's pixels, also with forced output channels; missing, truncated and refused files
    come back as sqoa_read / sqoa_write report them, without disturbing their neighbours."""
    rng = np.random.default_rng(77)
    shapes = [(64, 64, 4, "icon", 1), (64, 64, 4, "icon", 0), (300, 200, 3, "photo", 0), (257, 129, 4, "mixed", 1),
              (1000, 700, 3, "screen", 1), (31, 7, 4, "photo", 0), (640, 480, 4, "photo", 1), (5, 1, 3, "icon", 0),
              (120, 90, 1, "photo", 0), (77, 33, 2, "mixed", 0), (1920, 1080, 4, "mixed", 0), (2048, 1024, 3, "photo", 1)]
    imgs, descs, names = [], [], []
    for k, (w, h, c, kind, q) in enumerate(shapes):
        base = synth.image(kind, w, h, 4 if c in (2, 4) else 3, seed=int(rng.integers(1, 1000))).reshape(w * h, -1)
        if c == 1:
            im = base[:, 1].copy()
        elif c == 2:
            im = base[:, [1, 3]].copy()
        else:
            im = base
        imgs.append(np.ascontiguousarray(im).reshape(-1))
        descs.append(sb.Desc(w, h, c, 0, q))
        names.append(str(tmp_path / f"img{k}.{'qoi' if q else 'sqoa'}"))
    # one refused image (mono in QOI format, seqoia.h:477-480) in the middle
    imgs.insert(4, np.zeros(10 * 10, np.uint8))
    descs.insert(4, sb.Desc(10, 10, 1, 0, 1))
    names.insert(4, str(tmp_path / "refused.qoi"))
    sizes = sb.write_many(names, imgs, descs)
    want = [cpu.encode(im, d.width, d.height, d.channels, 0, d.qoi_compat) for im, d in zip(imgs, descs)]
    for k, (nm, w_) in enumerate(zip(names, want)):
        data = open(nm, "rb").read()
        if w_ is None:
            assert sizes[k] == 0 and data == b"", k
        else:
            assert sizes[k] == len(w_) and data == w_, (k, sizes[k], len(w_))
    # read back: plus a missing file and a truncated one
    trunc = str(tmp_path / "trunc.sqoa")
    open(trunc, "wb").write(want[2][:20])
    rnames = names + [str(tmp_path / "missing.sqoa"), trunc]
    for channels in (0, 4, 3):
        got = sb.read_many(rnames, channels)
        assert len(got) == len(rnames)
        for k, nm in enumerate(rnames):
            px, d = got[k]
            ref_px, ref_d = sb.read(nm, channels)
            if k >= len(names) or want[k] is None:
                assert px is None and ref_px is None, (k, channels)
                continue
            exp = cpu.decode(want[k], channels)[0]
            assert px is not None and np.array_equal(px, exp), (k, channels)
            assert (d.width, d.height, d.channels, d.qoi_compat) == (ref_d.width, ref_d.height, ref_d.channels, ref_d.qoi_compat)


@pytest.mark.gpu
def test_host_entry_points_from_several_threads(torch_cuda, cpu):
    """sqoa_encode / sqoa_decode called from four threads at once (the reference is called one image per core by
    sqoabench's totals): the library hands concurrent callers separate contexts (SQOA_B200_HOST_CONTEXTS, default 2) and
    queues the rest; every stream and every pixel buffer as the reference's."""
    import threading

    shapes = [(1920, 1080, 4, "mixed"), (2048, 1500, 3, "photo"), (640, 480, 4, "icon"), (3000, 1200, 3, "screen")]
    imgs = [synth.image(kind, w, h, c, seed=60 + i).reshape(-1) for i, (w, h, c, kind) in enumerate(shapes)]
    want = [[cpu.encode(im, w, h, c, 0, q) for q in (0, 1)] for im, (w, h, c, _k) in zip(imgs, shapes)]
    errors = []

    def work(k):
        try:
            for rep in range(3):
                i = (k + rep) % len(shapes)
                w, h, c, _kind = shapes[i]
                for q in (0, 1):
                    s = sb.encode(imgs[i], w, h, c, 0, q)
                    if s != want[i][q]:
                        errors.append(("encode", k, i, q))
                    px, _d = sb.decode(want[i][q], 0)
                    if px is None or not np.array_equal(px, imgs[i]):
                        errors.append(("decode", k, i, q))
        except Exception as e:  # noqa: BLE001
            errors.append(("exception", k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_transcode_batch_equals_reference_reencode(torch_cuda, cpu):
    """sqoa_b200_transcode_batch_device: every new stream == reference sqoa_encode(sqoa_decode(stream)), both directions,
    mixed shapes and channel counts, groups smaller than the batch (sqoaconv.c:65-84)."""
    torch = torch_cuda
    rng = np.random.default_rng(11)
    shapes = [(64, 64, 4, "icon"), (300, 200, 3, "photo"), (257, 129, 4, "mixed"), (1000, 700, 3, "screen"), (31, 7, 4, "photo"),
              (640, 480, 4, "photo"), (5, 1, 3, "icon")]
    imgs = [synth.image(kind, w, h, c, seed=int(rng.integers(1, 1000))) for w, h, c, kind in shapes]
    ctx = sb.Context(0)
    sp = torch.cuda.current_stream().cuda_stream
    al = lambda v: (v + 63) // 64 * 64
    for src, dst in ((0, 1), (1, 0)):
        streams = [cpu.encode(im, w, h, c, 0, src) for im, (w, h, c, _k) in zip(imgs, shapes)]
        offs, total = [], 0
        for (w, h, c, _k) in shapes:
            offs.append(total)
            total += al(sb.max_stream_size(w, h, c))
        host = np.zeros(total, dtype=np.uint8)
        for o, s in zip(offs, streams):
            host[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        d_src = torch.from_numpy(host).cuda()
        d_dst = torch.zeros(total, dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(len(shapes), dtype=torch.int32, device="cuda")
        d_st = torch.zeros(len(shapes), dtype=torch.int32, device="cuda")
        plan = ctx.transcode_plan([sb.Item(offs[i], offs[i], w, h, len(streams[i]), c, 0, src, 0)
                                   for i, (w, h, c, _k) in enumerate(shapes)], dst)
        for _ in range(2):
            ctx.transcode_batch(plan, d_src, d_dst, d_len, d_st, sp)
        torch.cuda.synchronize()
        assert int(d_st.abs().sum().item()) == 0
        out = d_dst.cpu().numpy()
        lens = d_len.cpu().numpy()
        for i, (im, (w, h, c, _k)) in enumerate(zip(imgs, shapes)):
            want = cpu.encode(im, w, h, c, 0, dst)
            assert out[offs[i]: offs[i] + lens[i]].tobytes() == want, (src, dst, i)


@pytest.mark.parametrize("qoi", [0, 1])
def test_sharded_encode_in_one_call_on_one_gpu(torch_cuda, cpu, qoi):
    """sqoa_b200_encode_sharded_device with a loop-back communicator: three shards of one image encoded one after the
    other on one GPU, the all-gather served from summaries computed beforehand; segments concatenated == reference."""
    torch = torch_cuda
    import ctypes as C

    w, h, ch = 1500, 601, 4
    img = synth.image("mixed", w, h, ch, seed=3)
    want = cpu.encode(img, w, h, ch, 0, qoi)
    ctx = sb.Context(0)
    sp = torch.cuda.current_stream().cuda_stream
    world = 3
    rows = [0, 200, 420, h]
    shards = [torch.from_numpy(img[rows[r]:rows[r + 1]].reshape(-1)).cuda() for r in range(world)]
    d_sums = torch.zeros(world * 80, dtype=torch.int32, device="cuda")
    for r in range(world):
        ctx.shard_summary(shards[r], (rows[r + 1] - rows[r]) * w, ch, qoi, d_sums[80 * r:], sp)
    torch.cuda.synchronize()

    def allgather(_user, d_send, d_recv, nbytes, _stream):  # every "rank" receives all three summaries
        assert nbytes == 320
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(d_recv), C.c_void_p(d_sums.data_ptr()), C.c_size_t(world * 320), C.c_int(3))
        return 0

    cb = sb.ALLGATHER_FN(allgather)
    out = b""
    for r in range(world):
        n_px = (rows[r + 1] - rows[r]) * w
        cap = n_px * 5 + 64
        d_seg = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(4, dtype=torch.int32, device="cuda")
        ctx.encode_sharded(sb.Comm(r, world, cb, None), shards[r], n_px, sb.Desc(w, h, ch, 0, qoi), d_seg, cap, d_len, sp)
        torch.cuda.synchronize()
        out += bytes(d_seg[: int(d_len[0].item())].cpu().numpy())
    assert out == want, first_difference(out, want)


@pytest.mark.gpu
@pytest.mark.parametrize("ch", [3, 4])
def test_sharded_decode_in_one_call_on_one_gpu(torch_cuda, cpu, ch):
    """sqoa_b200_decode_sharded_device with a loop-back communicator: the byte ranges of one SQOA stream decoded one
    after the other on one GPU, each in ONE call (three passes, device folds); the two all-gathers are served from
    summaries computed beforehand through the step-by-step API.  Pieces put together == the reference's pixels; a pixel
    buffer that is too small is reported through the status word with the size that is needed."""
    torch = torch_cuda
    import ctypes as C

    from seqoia_b200 import dist as sdist

    w, h = 1531, 420
    img = synth.image("mixed", w, h, ch, seed=6, cell=(61, 23))
    stream = np.frombuffer(cpu.encode(img, w, h, ch, 0, 0), dtype=np.uint8)
    want = img.reshape(-1)
    body_len = len(stream) - 15 - 8
    ctx = sb.Context(0)
    desc = sb.Desc(w, h, ch, 0, 0)
    sp = torch.cuda.current_stream().cuda_stream
    rt = C.cdll.LoadLibrary("libcudart.so.12")
    for world in (1, 3):
        cuts = sdist.stream_cuts(body_len, world)
        bufs, carries = [], []
        for r in range(world):
            b0, b1 = cuts[r], cuts[r + 1]
            tail = stream[15 + b0: min(len(stream), 15 + b1 + 32)]
            bufs.append((torch.from_numpy(tail.copy()).cuda(), len(tail), b1 - b0))
            carries.append(sb.DecCarry(sb.DEC_ENTRY, 0, 0, 0, 0, 1 if r == world - 1 else 0, b1 - b0, 0))
        gathered = {}
        d_sum = [torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(world)]
        for mode in (sb.DEC_ENTRY, sb.DEC_SCAN):  # what the ranks would exchange
            for r in range(world):
                carries[r].mode = mode
                ctx.decode_shard(bufs[r][0], bufs[r][1], desc, 0, carries[r], d_sum[r], None, 0, None, 0)
            torch.cuda.synchronize()
            gathered[mode] = torch.cat(d_sum).clone()
            sums = [sb.DecSummary.from_buffer_copy(x.cpu().numpy().tobytes()) for x in d_sum]
            for r in range(world):
                sb.fold_dec_carry(sums, r, carries[r])
        calls = [0]

        def allgather(_user, d_send, d_recv, nbytes, _stream):
            assert nbytes == 32
            src = gathered[sb.DEC_ENTRY if calls[0] % 2 == 0 else sb.DEC_SCAN]
            calls[0] += 1
            rt.cudaMemcpy(C.c_void_p(d_recv), C.c_void_p(src.data_ptr()), C.c_size_t(world * 32), C.c_int(3))
            return 0

        cb = sb.ALLGATHER_FN(allgather)
        d_info = torch.zeros(2, dtype=torch.int64, device="cuda")
        d_status = torch.ones(1, dtype=torch.int32, device="cuda")
        pieces, at = [], 0
        for r in range(world):
            d_px = torch.zeros(w * h * ch + 64, dtype=torch.uint8, device="cuda")
            ctx.decode_sharded(sb.Comm(r, world, cb, None), bufs[r][0], bufs[r][1], bufs[r][2], desc, 0, d_px, d_px.numel(),
                               d_info, d_status, sp)
            torch.cuda.synchronize()
            assert int(d_status.item()) == 0, (world, r)
            first, count = (int(v) for v in d_info.tolist())
            assert first == at and first == carries[r].pos, (world, r, first, at)
            at += count
            pieces.append(d_px[: count * ch].cpu().numpy())
            if r == world - 1:  # too small by one pixel: reported, with what is needed
                ctx.decode_sharded(sb.Comm(r, world, cb, None), bufs[r][0], bufs[r][1], bufs[r][2], desc, 0, d_px,
                                   (count - 1) * ch, d_info, d_status, sp)
                torch.cuda.synchronize()
                assert int(d_status.item()) == sb.E_CAPACITY
                assert int(d_info[1].item()) == count
        assert at == w * h
        assert np.array_equal(np.concatenate(pieces), want), world


@pytest.mark.gpu
@pytest.mark.parametrize("ch", [3, 4])
def test_sharded_qoi_decode_in_one_call_on_one_gpu(torch_cuda, cpu, ch):
    """sqoa_b200_decode_sharded_device on a QOI stream with a loop-back communicator: the byte ranges decoded one after
    the other on one GPU, each call handing its 544-byte carry to the next through the all-gather callback.  Pieces put
    together == the reference's pixels; a pixel buffer that is too small is reported; a stream the optimistic QOI
    decoder flags (half-transparent palette) is reported as not shardable."""
    torch = torch_cuda
    import ctypes as C

    from seqoia_b200 import dist as sdist

    rt = C.cdll.LoadLibrary("libcudart.so.12")
    ctx = sb.Context(0)
    sp = torch.cuda.current_stream().cuda_stream

    def run(stream, w, h, world, oc, capacity_px=None):
        body_len = len(stream) - 14 - 8
        cuts = sdist.stream_cuts(body_len, world)
        desc = sb.Desc(w, h, ch, 0, 1)
        carries = torch.zeros(64 * 544, dtype=torch.uint8, device="cuda")
        cur = [0]

        def allgather(_user, d_send, d_recv, nbytes, _stream):
            assert nbytes == 544
            rt.cudaMemcpy(C.c_void_p(carries.data_ptr() + 544 * cur[0]), C.c_void_p(d_send), C.c_size_t(544), C.c_int(3))
            rt.cudaMemcpy(C.c_void_p(d_recv), C.c_void_p(carries.data_ptr()), C.c_size_t(world * 544), C.c_int(3))
            return 0

        cb = sb.ALLGATHER_FN(allgather)
        raw = np.frombuffer(stream, dtype=np.uint8)
        pieces, verdicts = [], []
        for r in range(world):
            cur[0] = r
            b0, b1 = cuts[r], cuts[r + 1]
            avail = b1 - b0 + (8 if r == world - 1 else 64)
            buf = np.zeros(avail + 64, dtype=np.uint8)
            part = raw[14 + b0: 14 + b0 + avail]
            buf[: len(part)] = part
            d_body = torch.from_numpy(buf).cuda()
            cap = (w * h if capacity_px is None else capacity_px)
            d_px = torch.zeros(cap * oc + 64, dtype=torch.uint8, device="cuda")
            d_info = torch.zeros(2, dtype=torch.int64, device="cuda")
            d_status = torch.ones(1, dtype=torch.int32, device="cuda")
            ctx.decode_sharded(sb.Comm(r, world, cb, None), d_body, avail, b1 - b0, desc, oc, d_px, cap * oc, d_info, d_status, sp)
            torch.cuda.synchronize()
            first, count = (int(v) for v in d_info.tolist())
            verdicts.append((int(d_status.item()), first, count))
            pieces.append(d_px[: min(count, cap) * oc].cpu().numpy())
        return pieces, verdicts

    w, h = 1531, 420
    img = synth.image("mixed", w, h, ch, seed=9, cell=(61, 23))
    stream = cpu.encode(img, w, h, ch, 0, 1)
    for world, oc in ((1, ch), (3, ch), (4, 7 - ch)):
        want = cpu.decode(stream, oc)[0]
        pieces, verdicts = run(stream, w, h, world, oc)
        assert all(v[0] == 0 for v in verdicts), (world, verdicts)
        assert verdicts[0][1] == 0 and sum(v[2] for v in verdicts) == w * h
        assert all(verdicts[k][1] == verdicts[k - 1][1] + verdicts[k - 1][2] for k in range(1, world))
        assert np.array_equal(np.concatenate(pieces), want), world
    _, tight = run(stream, w, h, 3, ch, capacity_px=w * h // 10)
    assert any(v[0] == sb.E_CAPACITY and v[2] > w * h // 10 for v in tight), tight
    # the ordinary QOI decode on the same context afterwards
    px, _d = sb.decode(stream, ch)
    assert np.array_equal(px, cpu.decode(stream, ch)[0])
    if ch == 4:
        rng = np.random.default_rng(5)
        pal = rng.integers(0, 256, (40, 3), dtype=np.uint8)
        im = np.zeros((w * h, 4), np.uint8)
        im[:, :3] = pal[rng.integers(0, 40, w * h)]
        fresh = rng.random(w * h) < 0.3
        im[fresh, :3] = rng.integers(0, 256, (int(fresh.sum()), 3))
        im[:, 3] = 128
        s2 = cpu.encode(im.reshape(-1), w, h, 4, 0, 1)
        _, verdicts = run(s2, w, h, 3, 4)
        assert any(v[0] == sb.E_STREAM for v in verdicts), verdicts
        px, _d = sb.decode(s2, 4)  # one GPU decodes it (second attempt of the rows kernel)
        assert np.array_equal(px, im.reshape(-1))


@pytest.mark.gpu
def test_decode_shard_checks_its_pixel_buffer(torch_cuda, cpu):
    torch = torch_cuda
    w, h, ch = 400, 300, 4
    img = synth.image("photo", w, h, ch, seed=2)
    s = cpu.encode(img, w, h, ch, 0, 0)
    body = np.frombuffer(s, dtype=np.uint8)[15:-8]
    d_body = torch.from_numpy(np.concatenate([body, np.zeros(64, dtype=np.uint8)])).cuda()
    ctx = sb.Context(0)
    carry = sb.DecCarry(sb.DEC_PIXELS, 0, 0, 0, 0xff000000, 1, len(body), 0)
    small = torch.zeros(w * h * ch - 4, dtype=torch.uint8, device="cuda")
    with pytest.raises(sb.SqoaError):
        ctx.decode_shard(d_body, len(body) + 32, sb.Desc(w, h, ch, 0, 0), 0, carry, None, small, small.numel(), None, 0)
