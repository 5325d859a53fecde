"""BASELINE.json configs at FULL size against digests of the reference's own streams (tests/golden/digests_full.json,
made by oracle/make_golden_full.py from the compiled reference): the 20000x19999 RGBA image of cfg4 (the size cap of
seqoia.h:470, stream offsets up to 1.99e9 -- the `int` contract of seqoia.h:487-489), the whole 100,000-icon batch of
cfg3, the mixed corpus of cfg5, and forced 3 <-> 4 channel decodes (seqoia.h:790-805) of the full-size cfg2 streams."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
import seqoia_b200 as sb
from seqoia_b200 import synth
from util import golden


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a B200; there is no CPU fallback to test"
    torch.cuda.set_device(0)
    return torch


def _sha_of_device(t, n):
    h = hashlib.sha256()
    step = 256 << 20
    for a in range(0, n, step):
        h.update(t[a:min(n, a + step)].cpu().numpy().data)
    return h.hexdigest()


def test_cfg4_full_size_both_formats(torch_cuda):
    torch = torch_cuda
    w, h = 20000, 19999
    dig = golden("digests_full.json")["digests"]
    img = synth.cfg4(w, h)
    n_raw = w * h * 4
    d_px = torch.from_numpy(img.reshape(-1)).cuda()
    assert hashlib.sha256(img.reshape(-1).data).hexdigest() == dig[f"cfg4_{w}x{h}_rgba_q0"]["pixels_sha256"], "generator drifted"
    del img
    ctx = sb.Context(0)
    cap = sb.max_stream_size(w, h, 4)
    assert cap == 1_999_900_023
    d_s = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
    d_back = torch.empty(n_raw + 64, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(4, dtype=torch.int32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    for q in (0, 1):
        want = dig[f"cfg4_{w}x{h}_rgba_q{q}"]
        ctx.encode_device(d_px, sb.Desc(w, h, 4, 0, q), d_s, cap + 64, d_n, sp)
        torch.cuda.synchronize()
        n = int(d_n[0].item())
        assert n == want["stream_len"]
        assert _sha_of_device(d_s, n) == want["stream_sha256"]
        rc, dd, nbytes = sb.probe(bytes(d_s[:15].cpu().numpy()), n, 0)
        assert rc == sb.OK and nbytes == n_raw
        d_back.zero_()
        ctx.decode_device(d_s, n, dd, 0, d_back, n_raw, d_st, sp)
        torch.cuda.synchronize()
        assert int(d_st[0].item()) == 0
        assert bool(torch.equal(d_back[:n_raw], d_px))


@pytest.mark.parametrize("qoi", [0, 1])
def test_cfg3_all_100k_icons(torch_cuda, qoi):
    torch = torch_cuda
    n = 100_000
    want = golden("digests_full.json")["digests"][f"cfg3_icons_0_{n - 1}_q{qoi}"]
    icons = synth.cfg3(n)
    assert hashlib.sha256(icons.reshape(-1).data).hexdigest() == want["pixels_sha256"], "generator drifted"
    px_bytes = 64 * 64 * 4
    cap = (sb.max_stream_size(64, 64, 4) + 63) // 64 * 64
    ctx = sb.Context(0)
    sp = torch.cuda.current_stream().cuda_stream
    d_px = torch.from_numpy(icons.reshape(-1)).cuda()
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    plan = ctx.plan([sb.Item(i * px_bytes, i * cap, 64, 64, 0, 4, 0, qoi, 0) for i in range(n)])
    ctx.encode_batch(plan, d_px, d_out, d_len, sp)
    torch.cuda.synchronize()
    lens = d_len.cpu().numpy().astype(np.int64)
    assert int(lens.sum()) == want["stream_len"]
    out = d_out.cpu().numpy().reshape(n, cap)
    hsh = hashlib.sha256()
    for i in range(n):
        hsh.update(out[i, : lens[i]].data)
    assert hsh.hexdigest() == want["stream_sha256"]
    # and back
    dplan = ctx.plan([sb.Item(i * cap, i * px_bytes, 64, 64, int(lens[i]), 4, 0, qoi, 4) for i in range(n)], decode_=True)
    d_back = torch.zeros(n * px_bytes, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.decode_batch(dplan, d_out, d_back, d_status, sp)
    torch.cuda.synchronize()
    assert int(d_status.abs().sum().item()) == 0 and bool(torch.equal(d_back, d_px))


def test_cfg5_quarter_corpus(torch_cuda):
    torch = torch_cuda
    scale = 0.25
    dig = golden("digests_full.json")["digests"]
    shapes = synth.cfg5_shapes(scale)
    al = lambda v: (v + 63) // 64 * 64
    px_off, st_off, px_total, st_total = [], [], 0, 0
    for _k, w, h, c, _s in shapes:
        px_off.append(px_total)
        st_off.append(st_total)
        px_total += al(w * h * c)
        st_total += al(sb.max_stream_size(w, h, c))
    host = np.zeros(px_total, dtype=np.uint8)
    for (kind, w, h, c, seed), o in zip(shapes, px_off):
        synth.image(kind, w, h, c, seed=seed, out=host[o:o + w * h * c].reshape(h, w, c))
    ctx = sb.Context(0)
    sp = torch.cuda.current_stream().cuda_stream
    d_px = torch.from_numpy(host).cuda()
    for q in (0, 1):
        want = dig[f"cfg5_scale{scale}_q{q}"]
        assert want["n"] == len(shapes)
        d_st = torch.zeros(st_total, dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(len(shapes), dtype=torch.int32, device="cuda")
        plan = ctx.plan([sb.Item(px_off[i], st_off[i], w, h, 0, c, 0, q, 0) for i, (_k, w, h, c, _s) in enumerate(shapes)])
        ctx.encode_batch(plan, d_px, d_st, d_len, sp)
        torch.cuda.synchronize()
        lens = d_len.cpu().numpy().astype(np.int64)
        assert int(lens.sum()) == want["stream_len"]
        out = d_st.cpu().numpy()
        hsh = hashlib.sha256()
        for i in range(len(shapes)):
            hsh.update(out[st_off[i]: st_off[i] + lens[i]].data)
        assert hsh.hexdigest() == want["stream_sha256"]


@pytest.mark.parametrize("src_ch,out_ch", [(3, 4), (4, 3)])
@pytest.mark.parametrize("qoi", [0, 1])
def test_cfg2_forced_channel_conversion(torch_cuda, src_ch, out_ch, qoi):
    """Decode the full-size cfg2 stream into the other pixel size (seqoia.h:790-805) and compare with the reference."""
    cpu = oracle.best()
    img = synth.cfg2(channels=src_ch)
    s = sb.encode(img, 3840, 2160, src_ch, 0, qoi)
    px, d = sb.decode(s, out_ch)
    ref_px, ref_d = cpu.decode(s, out_ch)
    assert px is not None and px.size == 3840 * 2160 * out_ch
    assert np.array_equal(px, ref_px)
    assert (d.width, d.height, d.channels, d.qoi_compat) == (ref_d.width, ref_d.height, ref_d.channels, ref_d.qoi_compat)
