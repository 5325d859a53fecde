#!/usr/bin/env python
"""bench.py -- SQOA / QOI encode + decode throughput on B200 (BASELINE.json metric).

One "step" is one pass of the hot path over BASELINE.json configs[1]: the
3840x2160 RGB photo-like synthetic image, in both formats -- SQOA encode, SQOA
decode, QOI encode, QOI decode (4 x 8,294,400 pixels).  `value` is pixels
processed per second over the whole step with every buffer already resident in
HBM; `e2e` is the same step through the reference's own entry points
(sqoa_encode / sqoa_decode on HOST buffers, copies inside the timed region).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg3]

Under torchrun (N > 1) every rank runs the step on its own images (independent
units, no data-path collective: weak scaling); the time is the max over ranks.

L2 hygiene: the step rotates over REPLICAS distinct copies of every buffer
(> 3x the 126 MB L2 in total) and decodes a stream that was encoded a full
rotation earlier, so no leg finds its input in L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SQOA+QOI encode/decode throughput, 3840x2160 RGB (Mpx/s, device-resident)"
UNIT = "Mpx/s"
REPLICAS = 4


def ncu_traffic(leg: str):
    """DRAM bytes per launch sequence of `leg` from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return int(json.load(f)["legs"][leg]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------
def cpu_step_times(codec, img, w, h, ch, copies: int, threads: int):
    """Seconds for (sqoa enc, sqoa dec, qoi enc, qoi dec) of `copies` images on `threads` cores."""
    import ctypes as C

    import oracle

    drv = oracle.timing_driver()
    px = np.ascontiguousarray(np.broadcast_to(img.reshape(1, -1), (copies, img.size)))
    out = {}
    streams = {}
    for q, name in ((0, "sqoa"), (1, "qoi")):
        total = C.c_longlong(0)
        out[f"{name}_encode"] = drv.cb_time_encode(codec.enc_ptr, px.ctypes.data, img.size, copies, w, h, ch, q,
                                                   threads, C.byref(total))
        s = codec.encode(img, w, h, ch, 0, q)
        streams[name] = s
        blob = np.frombuffer(s * copies, dtype=np.uint8)
        offs = (C.c_longlong * copies)(*[i * len(s) for i in range(copies)])
        lens = (C.c_int * copies)(*[len(s)] * copies)
        npx = C.c_longlong(0)
        out[f"{name}_decode"] = drv.cb_time_decode(codec.dec_ptr, blob.ctypes.data, offs, lens, copies, 0, threads,
                                                   C.byref(npx))
        assert npx.value == copies * w * h
    return out, streams


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    from seqoia_b200 import synth

    oracle.build(with_reference=True)
    codec = oracle.best()
    w, h, ch = 3840, 2160, 3
    img = synth.cfg2()
    cores = os.cpu_count() or 1
    copies = cores  # one image per core
    per_step = []
    for i in range(args.warmup + args.steps):
        t, _ = cpu_step_times(codec, img, w, h, ch, copies, cores)
        if i >= args.warmup:
            per_step.append(sum(t.values()))
    secs = float(np.mean(per_step))
    value = 4 * copies * w * h / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg2: 3840x2160 RGB photo-like, SQOA+QOI encode+decode (4 legs per step)",
                   "images_per_step": copies},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": codec.kind,
                         "sample": f"{copies} copies of the cfg2 image per step, one image per core, "
                                   f"4 legs (sqoa/qoi x enc/dec), malloc/free inside the timed region like sqoabench"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch

    if world > 1:
        # N processes share this box's host cores: the library's copy threads (8 per context by default) are scaled
        # down so that the end-to-end leg of one rank does not starve the others
        os.environ.setdefault("SQOA_B200_COPY_THREADS", str(max(1, (os.cpu_count() or 16) // world - 1)))
    import seqoia_b200 as sb
    from seqoia_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    w, h, ch = 3840, 2160, 3
    npx = w * h
    img = synth.cfg2()  # every rank: the same recipe (independent units; weak scaling)
    ctx = sb.Context(local_rank)
    cap = sb.max_stream_size(w, h, ch)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    d_px = [torch.from_numpy(img.reshape(-1)).to(dev) for _ in range(REPLICAS)]
    d_stream = {q: [torch.empty(cap + 64, dtype=torch.uint8, device=dev) for _ in range(REPLICAS)] for q in (0, 1)}
    d_out = {q: [torch.empty(npx * ch + 64, dtype=torch.uint8, device=dev) for _ in range(REPLICAS)] for q in (0, 1)}
    d_len = {q: [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(REPLICAS)] for q in (0, 1)}
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    desc = {q: sb.Desc(w, h, ch, 0, q) for q in (0, 1)}

    # prime: every replica's streams exist before the timed region (decode reads them)
    for r in range(REPLICAS):
        for q in (0, 1):
            ctx.encode_device(d_px[r], desc[q], d_stream[q][r], cap, d_len[q][r], sptr)
    torch.cuda.synchronize()
    slen = {q: int(d_len[q][0][0].item()) for q in (0, 1)}
    ddesc = {}
    for q in (0, 1):
        hdr = bytes(d_stream[q][0][:15].cpu().numpy())
        rc, dd, nbytes = sb.probe(hdr, slen[q], 0)
        assert rc == sb.OK and nbytes == npx * ch
        ddesc[q] = dd

    legs = ["sqoa_encode", "sqoa_decode", "qoi_encode", "qoi_decode"]
    alg_bytes = {"sqoa_encode": npx * ch + slen[0], "sqoa_decode": npx * ch + slen[0],
                 "qoi_encode": npx * ch + slen[1], "qoi_decode": npx * ch + slen[1]}
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]

    def step(i, events=None):
        r, rd = i % REPLICAS, (i + 1) % REPLICAS
        if events:
            events[0].record(stream)
        ctx.encode_device(d_px[r], desc[0], d_stream[0][r], cap, d_len[0][r], sptr)
        if events:
            events[1].record(stream)
        ctx.decode_device(d_stream[0][rd], slen[0], ddesc[0], 0, d_out[0][rd], npx * ch, d_status, sptr)
        if events:
            events[2].record(stream)
        ctx.encode_device(d_px[rd], desc[1], d_stream[1][r], cap, d_len[1][r], sptr)
        if events:
            events[3].record(stream)
        ctx.decode_device(d_stream[1][rd], slen[1], ddesc[1], 0, d_out[1][rd], npx * ch, d_status, sptr)
        if events:
            events[4].record(stream)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, ev[i])
    t_end.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    if dist:
        tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
        dist.barrier()
    leg_ms = {name: float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(args.steps)]))
              for k, name in enumerate(legs)}

    # parity spot check on the timed buffers (not timed): decoded pixels == input
    ok = all(bool(torch.equal(d_out[q][r][: npx * ch], d_px[r])) for q in (0, 1) for r in range(REPLICAS))

    # ---- e2e: the reference's own entry points on host buffers -------------------------
    # sqoa_encode / sqoa_decode called through the C ABI exactly as a C program would: the input
    # pixels sit in pinned host memory, results are the library's malloc() buffers (freed here).
    import ctypes as C

    L = sb.lib()
    host_px = torch.from_numpy(img.reshape(-1).copy()).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    h2d = d2h = 0
    e2e_secs = []
    e2e_ok = True
    for i in range(1 + e2e_steps):
        t0 = time.perf_counter()
        h2d = d2h = 0
        for q in (0, 1):
            d = sb.Desc(w, h, ch, 0, q)
            n = C.c_int(0)
            sp = L.sqoa_encode(host_px.data_ptr(), C.byref(d), C.byref(n))
            d2 = sb.Desc()
            pp = L.sqoa_decode(sp, n.value, C.byref(d2), 0)
            h2d += host_px.numel() + n.value
            d2h += n.value + npx * ch
            if i == 0:  # parity of the e2e path, outside the timed iterations
                back = np.frombuffer(C.string_at(pp, npx * ch), dtype=np.uint8)
                e2e_ok = e2e_ok and bool(np.array_equal(back, img.reshape(-1))) and n.value == slen[q]
            L._free(sp)
            L._free(pp)
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_secs.append(dt)
    ok = ok and e2e_ok
    e2e_s = float(np.mean(e2e_secs))
    if dist:
        tm = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())

    if rank != 0:
        return
    peak, peak_kind = measured_hbm_peak()
    ms_per_step = total_ms / args.steps
    value = world * 4 * npx / (ms_per_step * 1e-3) / 1e6
    dom = max(legs, key=lambda k: leg_ms[k])
    leg_report = {}
    for k in legs:
        gbs = alg_bytes[k] / (leg_ms[k] * 1e-3) / 1e9
        leg_report[k] = {"ms": leg_ms[k], "mpx_s": npx / (leg_ms[k] * 1e-3) / 1e6, "gb_s": gbs,
                         "frac_of_measured_hbm": gbs / peak, "frac_of_nominal_8tbs": gbs / 8000.0,
                         "algorithmic_bytes": alg_bytes[k]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg2: 3840x2160 RGB photo-like, SQOA+QOI encode+decode (4 legs per step)",
                   "l2": f"{REPLICAS} rotating buffer replicas (>3x L2); decode reads a stream encoded a rotation earlier",
                   "parallelism": f"independent images, {world} GPU(s), no collective"},
        "legs": leg_report,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": leg_report[dom]["gb_s"], "peak": peak,
                     "unit": "GB/s", "frac": leg_report[dom]["gb_s"] / peak, "traffic": ncu_traffic(dom),
                     "traffic_note": "dram read+write bytes per launch sequence, ncu capture in profiles/r01_traffic.json",
                     "peak_kind": f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json)",
                     "per_leg_frac": {k: leg_report[k]["frac_of_measured_hbm"] for k in legs}},
        "e2e": {"value": world * 4 * npx / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "sqoa_encode + sqoa_decode on host buffers, both formats"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_spot_check": ok,
    }
    # cpu_baseline: the reference on this box's host cores, bounded sample, rank 0 only, N == 1 only
    if world == 1:
        try:
            import oracle

            oracle.build(with_reference=True)
            codec = oracle.best()
            t1, _ = cpu_step_times(codec, img, w, h, ch, 1, 1)
            cores = os.cpu_count() or 1
            tn, _ = cpu_step_times(codec, img, w, h, ch, cores, cores)
            line["cpu_baseline"] = {
                "value": 4 * npx / sum(t1.values()) / 1e6, "unit": UNIT, "cores": 1, "kind": codec.kind,
                "sample": "one cfg2 image, 4 legs, single thread",
                "legs_mpx_s": {k: npx / v / 1e6 for k, v in t1.items()},
                "all_cores": {"value": 4 * cores * npx / sum(tn.values()) / 1e6, "cores": cores,
                              "sample": f"{cores} copies, one image per core"},
            }
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# other BASELINE.json configs (not the driver's default line): --workload cfg3 | cfg4
# ---------------------------------------------------------------------------------------
def run_cfg3(args, rank, world, local_rank):
    """configs[2]: 100k 64x64 RGBA icons sharded by image index (no collective): encode + decode, both formats."""
    import torch

    import seqoia_b200 as sb
    from seqoia_b200 import dist as sdist
    from seqoia_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    n_total = args.images
    lo, hi = sdist.shard_range(n_total, world, rank)
    n = hi - lo
    icons = synth.cfg3(n, first=lo)
    px_bytes = 64 * 64 * 4
    cap = (sb.max_stream_size(64, 64, 4) + 63) // 64 * 64
    ctx = sb.Context(local_rank)
    sptr = torch.cuda.current_stream().cuda_stream
    d_px = torch.from_numpy(icons.reshape(-1)).to(dev)
    d_out = {q: torch.empty(n * cap, dtype=torch.uint8, device=dev) for q in (0, 1)}
    d_len = {q: torch.zeros(n, dtype=torch.int32, device=dev) for q in (0, 1)}
    d_back = torch.empty(n * px_bytes, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    enc_plan = {q: ctx.plan([sb.Item(i * px_bytes, i * cap, 64, 64, 0, 4, 0, q, 0) for i in range(n)]) for q in (0, 1)}
    for q in (0, 1):
        ctx.encode_batch(enc_plan[q], d_px, d_out[q], d_len[q], sptr)
    torch.cuda.synchronize()
    lens = {q: d_len[q].cpu().numpy() for q in (0, 1)}
    dec_plan = {q: ctx.plan([sb.Item(i * cap, i * px_bytes, 64, 64, int(lens[q][i]), 4, 0, q, 4) for i in range(n)],
                            decode_=True) for q in (0, 1)}
    legs = ["sqoa_encode", "sqoa_decode", "qoi_encode", "qoi_decode"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    stream = torch.cuda.current_stream()

    def step(events=None):
        for k, (q, dec) in enumerate(((0, False), (0, True), (1, False), (1, True))):
            if events:
                events[k].record(stream)
            if dec:
                ctx.decode_batch(dec_plan[q], d_out[q], d_back, d_status, sptr)
            else:
                ctx.encode_batch(enc_plan[q], d_px, d_out[q], d_len[q], sptr)
        if events:
            events[4].record(stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(ev[i])
    t1.record(stream)
    torch.cuda.synchronize()
    total_ms = t0.elapsed_time(t1)
    ok = bool(torch.equal(d_back, d_px)) and int(d_status.abs().sum().item()) == 0
    if dist:
        tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
    if rank != 0:
        return
    peak, peak_kind = measured_hbm_peak()
    leg_ms = {name: float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(args.steps)]))
              for k, name in enumerate(legs)}
    npx = n * 4096
    rep = {}
    for k in legs:
        q = 0 if k.startswith("sqoa") else 1
        b = n * px_bytes + int(lens[q].sum())
        rep[k] = {"ms": leg_ms[k], "mpx_s": npx / (leg_ms[k] * 1e-3) / 1e6, "gb_s": b / (leg_ms[k] * 1e-3) / 1e9,
                  "frac_of_measured_hbm": b / (leg_ms[k] * 1e-3) / 1e9 / peak, "algorithmic_bytes": b}
    ms = total_ms / args.steps
    print(json.dumps({
        "metric": "SQOA+QOI encode/decode throughput, 64x64 RGBA icon batch (Mpx/s, device-resident)",
        "value": 4 * n_total * 4096 / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"cfg3: {n_total} 64x64 RGBA icons sharded by image index over {world} GPU(s), 4 legs",
                   "l2": "per-GPU working set > 3 GB, far larger than L2", "images_per_gpu": n},
        "legs_rank0": rep, "gpu_launches": int(ctx.launches - launches0), "parity_spot_check": ok}), flush=True)


def run_cfg5(args, rank, world, local_rank):
    """configs[4]: mixed-size corpus mirroring the qoi test-suite mix; SQOA<->QOI transcode on the device (the
    pixels never leave HBM).  Images shard over the GPUs by index, balanced by pixel count; no collective."""
    import torch

    import seqoia_b200 as sb
    from seqoia_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    shapes = synth.cfg5_shapes(args.scale)
    # contiguous index ranges with (nearly) equal pixel counts
    px_cum = np.cumsum([w * h for _k, w, h, _c, _s in shapes])
    cut = [int(np.searchsorted(px_cum, px_cum[-1] * r / world)) for r in range(world + 1)]
    cut[0], cut[-1] = 0, len(shapes)
    mine = shapes[cut[rank]:cut[rank + 1]]
    n = len(mine)
    al = lambda v: (v + 63) // 64 * 64
    px_off, st_off, px_total, st_total = [], [], 0, 0
    for _k, w, h, c, _s in mine:
        px_off.append(px_total)
        st_off.append(st_total)
        px_total += al(w * h * c)
        st_total += al(sb.max_stream_size(w, h, c))
    host = np.zeros(px_total, dtype=np.uint8)
    for (kind, w, h, c, seed), o in zip(mine, px_off):
        synth.image(kind, w, h, c, seed=seed, out=host[o:o + w * h * c].reshape(h, w, c))
    ctx = sb.Context(local_rank)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    d_px = torch.from_numpy(host).to(dev)
    d_mid = torch.zeros(px_total, dtype=torch.uint8, device=dev)  # decoded pixels, device only
    d_st = {q: torch.zeros(st_total, dtype=torch.uint8, device=dev) for q in (0, 1)}   # direct encodes (sources)
    d_tr = {q: torch.zeros(st_total, dtype=torch.uint8, device=dev) for q in (0, 1)}   # transcoded streams
    d_len = {q: torch.zeros(n, dtype=torch.int32, device=dev) for q in (0, 1)}
    d_len_tr = {q: torch.zeros(n, dtype=torch.int32, device=dev) for q in (0, 1)}
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    enc_plan = {q: ctx.plan([sb.Item(px_off[i], st_off[i], w, h, 0, c, 0, q, 0)
                             for i, (_k, w, h, c, _s) in enumerate(mine)]) for q in (0, 1)}
    for q in (0, 1):
        ctx.encode_batch(enc_plan[q], d_px, d_st[q], d_len[q], sptr)
    torch.cuda.synchronize()
    lens = {q: d_len[q].cpu().numpy() for q in (0, 1)}
    dec_plan = {q: ctx.plan([sb.Item(st_off[i], px_off[i], w, h, int(lens[q][i]), c, 0, q, c)
                             for i, (_k, w, h, c, _s) in enumerate(mine)], decode_=True) for q in (0, 1)}
    legs = ["sqoa_to_qoi", "qoi_to_sqoa"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

    def step(events=None):
        for k, (src, dst) in enumerate(((0, 1), (1, 0))):
            if events:
                events[k].record(stream)
            ctx.decode_batch(dec_plan[src], d_st[src], d_mid, d_status, sptr)
            ctx.encode_batch(enc_plan[dst], d_mid, d_tr[dst], d_len_tr[dst], sptr)
        if events:
            events[2].record(stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(ev[i])
    t1.record(stream)
    torch.cuda.synchronize()
    total_ms = t0.elapsed_time(t1)
    # a transcoded stream must equal the direct encoding of the original pixels, byte for byte
    ok = all(bool(torch.equal(d_tr[q], d_st[q])) and bool(torch.equal(d_len_tr[q], d_len[q])) for q in (0, 1))
    ok = ok and int(d_status.abs().sum().item()) == 0
    npx = int(sum(w * h for _k, w, h, _c, _s in mine))
    tot = torch.tensor([total_ms, float(npx), float(lens[0].sum() + lens[1].sum())], dtype=torch.float64, device=dev)
    if dist:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_ms = float(mx[0].item())
    if rank != 0:
        return
    peak, _kind = measured_hbm_peak()
    npx_all, bytes_all = float(tot[1].item()), float(tot[2].item())
    ms = total_ms / args.steps
    leg_ms = {name: float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(args.steps)]))
              for k, name in enumerate(legs)}
    rep = {k: {"ms": leg_ms[k], "mpx_s": npx / (leg_ms[k] * 1e-3) / 1e6,
               "gb_s": float(lens[0].sum() + lens[1].sum()) / (leg_ms[k] * 1e-3) / 1e9} for k in legs}
    print(json.dumps({
        "metric": "SQOA<->QOI transcode throughput, mixed corpus (Mpx/s, device-resident)",
        "value": 2 * npx_all / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"cfg5: {len(shapes)} images of the qoi-suite mix (scale {args.scale}), "
                               f"{npx_all / 1e6:.0f} Mpx, SQOA->QOI and QOI->SQOA on the device, {world} GPU(s)",
                   "l2": "working set far larger than L2", "bytes": "stream in + stream out per direction",
                   "stream_gb_s": 2 * bytes_all / (ms * 1e-3) / 1e9},
        "legs_rank0": rep, "gpu_launches": int(ctx.launches - launches0), "parity_spot_check": ok}), flush=True)


def run_cfg4(args, rank, world, local_rank):
    """configs[3]: one 20000x19999 RGBA image, scanline-sharded; only boundary summaries cross GPUs."""
    import torch

    import seqoia_b200 as sb
    from seqoia_b200 import dist as sdist
    from seqoia_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, h = args.width, args.height
    y0, y1 = sdist.shard_rows(h, world, rank)
    mine = synth.cfg4_rows(y0, y1, w, h)
    n_px = (y1 - y0) * w
    ctx = sb.Context(local_rank)
    sptr = torch.cuda.current_stream().cuda_stream
    stream = torch.cuda.current_stream()
    d_px = torch.from_numpy(mine.reshape(-1)).to(dev)
    cap = n_px * 5 + 64
    d_seg = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(4, dtype=torch.int32, device=dev)
    res = {}
    for q, name in ((0, "sqoa"), (1, "qoi")):
        desc = sb.Desc(w, h, 4, 0, q)
        times = []
        for i in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            sdist.encode_sharded_device(ctx, d_px, n_px, desc, d_seg, cap, d_len, None, sptr)
            t1.record(stream)
            torch.cuda.synchronize()
            if i >= args.warmup:
                times.append(t0.elapsed_time(t1))
        ms = float(np.mean(times))
        seg_len = int(d_len[0].item())
        tot = torch.tensor([ms, float(seg_len)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tot.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tot.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, stream_len = float(mx[0].item()), int(sm[1].item())
        else:
            stream_len = seg_len
        res[f"{name}_encode"] = {"ms": ms, "mpx_s": w * h / (ms * 1e-3) / 1e6, "stream_bytes": stream_len,
                                 "gb_s": (w * h * 4 + stream_len) / (ms * 1e-3) / 1e9}
        if world > 1 and q == 0:
            # stream-sharded SQOA decode: every rank decodes one byte range of the stream (cut on decoder tile
            # boundaries); only the 8-word shard summaries cross GPUs (two NCCL all-gathers).  Setup, not timed:
            # the encoder's row-sharded segments are gathered so that every rank can take its byte range.
            lens_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(lens_all, torch.tensor([seg_len], dtype=torch.int64, device=dev))
            lens_all = [int(x.item()) for x in lens_all]
            pad = max(lens_all)
            mine_seg = torch.zeros(pad, dtype=torch.uint8, device=dev)
            mine_seg[:seg_len] = d_seg[:seg_len]
            parts = [torch.empty(pad, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(parts, mine_seg)
            full = torch.cat([parts[r][: lens_all[r]] for r in range(world)] + [torch.zeros(64, dtype=torch.uint8, device=dev)])
            del parts, mine_seg
            total = sum(lens_all)
            cuts = sdist.stream_cuts(total - 23, world)
            b0, b1 = cuts[rank], cuts[rank + 1]
            d_body = full[15 + b0:]
            avail = min(total - (15 + b0), b1 - b0 + 32)
            d_sum = torch.zeros(8, dtype=torch.int32, device=dev)
            pool = {}

            def alloc(nbytes):
                if "buf" not in pool or pool["buf"].numel() < nbytes:
                    pool["buf"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                return pool["buf"]

            times = []
            for i in range(args.warmup + args.steps):
                torch.cuda.synchronize()
                dist.barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(stream)
                d_out, first_px, n_mine = sdist.decode_stream_shard(ctx, d_body, avail, b1 - b0, desc, 0, rank, world,
                                                                    d_sum, alloc, sptr)
                t1.record(stream)
                torch.cuda.synchronize()
                if i >= args.warmup:
                    times.append(t0.elapsed_time(t1))
            # parity: my pixel range against the same range of the synthetic image (regenerated on the host)
            ya, yb = first_px // w, min(h, (first_px + n_mine + w - 1) // w)
            ref = synth.cfg4_rows(ya, yb, w, h).reshape(-1)[(first_px - ya * w) * 4: (first_px - ya * w + n_mine) * 4]
            ok = bool(torch.equal(d_out[: n_mine * 4].cpu(), torch.from_numpy(ref.copy())))
            tot = torch.tensor([float(np.mean(times)), float(n_mine), float(ok)], dtype=torch.float64, device=dev)
            mx = tot.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tot.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms = float(mx[0].item())
            res["sqoa_decode"] = {"ms": ms, "mpx_s": w * h / (ms * 1e-3) / 1e6,
                                  "gb_s": (w * h * 4 + total) / (ms * 1e-3) / 1e9,
                                  "pixels_decoded": int(sm[1].item()), "round_trip_ok": int(sm[2].item()) == world,
                                  "passes": "entry + scan + pixels, two all-gathers of 32-byte summaries"}
            del full, d_body, pool
        if world == 1:  # single GPU: decode the whole stream back and byte-compare
            rc, dd, nbytes = sb.probe(bytes(d_seg[:15].cpu().numpy()), seg_len, 0)
            d_back = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
            d_st = torch.zeros(4, dtype=torch.int32, device=dev)
            times = []
            for i in range(args.warmup + args.steps):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(stream)
                ctx.decode_device(d_seg, seg_len, dd, 0, d_back, nbytes, d_st, sptr)
                t1.record(stream)
                torch.cuda.synchronize()
                if i >= args.warmup:
                    times.append(t0.elapsed_time(t1))
            ms = float(np.mean(times))
            res[f"{name}_decode"] = {"ms": ms, "mpx_s": w * h / (ms * 1e-3) / 1e6,
                                     "gb_s": (w * h * 4 + stream_len) / (ms * 1e-3) / 1e9,
                                     "round_trip_ok": bool(torch.equal(d_back[:nbytes], d_px))}
            del d_back
    if rank != 0:
        return
    peak, _ = measured_hbm_peak()
    for v in res.values():
        v["frac_of_measured_hbm"] = v["gb_s"] / peak
    total_ms = sum(v["ms"] for v in res.values())
    print(json.dumps({
        "metric": f"SQOA+QOI encode throughput, one {w}x{h} RGBA image scanline-sharded (Mpx/s, device-resident)",
        "value": len(res) * w * h / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"cfg4: one {w}x{h} RGBA image, rows sharded over {world} GPU(s); "
                               "exchange = all-gather of 320-byte shard summaries (NCCL)",
                   "l2": "inputs far larger than L2"},
        "legs": res, "gpu_launches": int(ctx.launches)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=0.25, help="cfg5: fraction of the 2,851-image corpus")
    ap.add_argument("--images", type=int, default=100_000, help="cfg3: images in the batch")
    ap.add_argument("--width", type=int, default=20000, help="cfg4")
    ap.add_argument("--height", type=int, default=19999, help="cfg4")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "cfg3":
        run_cfg3(args, rank, world, local_rank)
    elif args.workload == "cfg5":
        run_cfg5(args, rank, world, local_rank)
    elif args.workload == "cfg4":
        run_cfg4(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
