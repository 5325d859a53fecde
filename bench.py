#!/usr/bin/env python
"""bench.py -- SQOA / QOI encode + decode throughput on B200 (BASELINE.json metric).

Headline (the ONE JSON line's top-level keys): BASELINE.json configs[1], the 3840x2160 RGB photo-like synthetic
image in both formats.  One "step" is one pass of the hot path over a batch of IMAGES_PER_STEP such images (distinct
device buffers, far larger than L2): SQOA encode, SQOA decode, QOI encode, QOI decode -- four launches of the batch
entry points.  `value` is pixels processed per second over the whole step with every buffer already resident in HBM;
`e2e` is the same work through the reference's own entry points (sqoa_encode / sqoa_decode of seqoia.h:363,374) on
pageable HOST buffers, copies inside the timed region.

The other BASELINE.json configs ride in the same line under "configs": cfg1 (1920x1080 RGBA round trip, latency),
cfg3 (100k icons, sharded by image index), cfg4 (one 20000x19999 RGBA image; scanline-sharded with a boundary
exchange when N > 1) and cfg5 (SQOA <-> QOI transcode of the mixed corpus), each with per-leg GB/s against the
measured HBM peak of the GPUs used, byte-level parity against the reference's digests (tests/golden/digests_full.json)
and the reference seqoia.h timed on this box's host cores beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--only cfg2|cfg1|cfg3|cfg4|cfg5] [--skip-configs]

Under torchrun (N > 1) every rank runs the headline step on its own images (independent units, no data-path
collective: weak scaling); the time is the max over ranks.  cfg3 / cfg5 shard by image index (strong scaling, no
collective), cfg4 by scanlines (strong scaling, one all-gather of 320-byte boundary summaries).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import hashlib
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
IMAGES_PER_STEP = 16
WORKLOAD = "cfg2: 3840x2160 RGB photo-like, SQOA+QOI encode+decode (4 legs per image)"
# the same dict in both arms (the driver compares them)
CONFIG = {"workload": WORKLOAD}


def traffic_of(leg: str):
    """DRAM bytes per launch of `leg` from the committed ncu capture (profiles/r02_traffic.json), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return int(json.load(f)["legs"][leg]["dram_bytes_per_launch"]), name
        except Exception:
            continue
    return None, None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def golden_full():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "digests_full.json")) as f:
            return json.load(f)["digests"]
    except Exception:
        return {}


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
# the reference's own CPU implementation on the host cores (reference arm / cpu_baseline)
# ---------------------------------------------------------------------------------------
def cpu_codec():
    import oracle

    oracle.build(with_reference=True)
    return oracle.best()


def cpu_uniform_times(codec, img, w, h, ch, copies: int, threads: int):
    """Seconds for (sqoa enc, sqoa dec, qoi enc, qoi dec) of `copies` images of one shape on `threads` cores; malloc /
    free of the results inside the timed region, as in sqoabench.c:490-538."""
    import ctypes as C

    import oracle

    drv = oracle.timing_driver()
    flat = img.reshape(copies, -1) if img.ndim > 1 and img.shape[0] == copies and copies > 1 and img.size == copies * w * h * ch \
        else np.ascontiguousarray(np.broadcast_to(img.reshape(1, -1), (copies, img.size)))
    out = {}
    for q, name in ((0, "sqoa"), (1, "qoi")):
        total = C.c_longlong(0)
        out[f"{name}_encode"] = drv.cb_time_encode(codec.enc_ptr, flat.ctypes.data, flat.shape[1], copies, w, h, ch, q,
                                                   threads, C.byref(total))
        streams = [codec.encode(flat[i], w, h, ch, 0, q) for i in range(copies)] if flat.shape[0] == copies and img.size != flat.shape[1] \
            else [codec.encode(flat[0], w, h, ch, 0, q)] * copies
        blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
        lens_l = [len(s) for s in streams]
        offs = (C.c_longlong * copies)(*np.concatenate([[0], np.cumsum(lens_l)[:-1]]).astype(np.int64).tolist())
        lens = (C.c_int * copies)(*lens_l)
        npx = C.c_longlong(0)
        out[f"{name}_decode"] = drv.cb_time_decode(codec.dec_ptr, blob.ctypes.data, offs, lens, copies, 0, threads,
                                                   C.byref(npx))
        assert npx.value == copies * w * h
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    from seqoia_b200 import synth

    codec = cpu_codec()
    w, h, ch = 3840, 2160, 3
    img = synth.cfg2()
    cores = os.cpu_count() or 1
    copies = cores  # one image per core
    per_step = []
    for i in range(args.warmup + args.steps):
        t = cpu_uniform_times(codec, img, w, h, ch, copies, cores)
        if i >= args.warmup:
            per_step.append(sum(t.values()))
    secs = float(np.mean(per_step))
    value = 4 * copies * w * h / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": dict(CONFIG),
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
class Env:
    """device, rank and collective plumbing shared by the workloads"""

    def __init__(self, rank, world, local_rank):
        import torch

        if world > 1:
            # N processes share this box's host cores: the library's copy threads (8 per context by default) are
            # scaled down so that the end-to-end leg of one rank does not starve the others
            os.environ.setdefault("SQOA_B200_COPY_THREADS", str(max(2, (os.cpu_count() or 16) // world)))
        import seqoia_b200 as sb

        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
        self.torch, self.sb = torch, sb
        self.rank, self.world, self.local_rank = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            self.dist = dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = sb.Context(local_rank)
        # QOI decodes without the host looking at the device in between (sqoa_b200_ctx_set_qoi_nowait): every leg of
        # a step is then queued asynchronously.  SQOA_BENCH_QOI_NOWAIT=0 restores the default mode of the library.
        self.qoi_nowait = os.environ.get("SQOA_BENCH_QOI_NOWAIT", "1") != "0"
        self.ctx.set_qoi_nowait(self.qoi_nowait)
        self.stream = torch.cuda.current_stream()
        self.sptr = self.stream.cuda_stream
        self.peak, self.peak_kind = measured_hbm_peak()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        if not self.dist:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if not self.dist:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        return self.sum_over_ranks(1.0 if ok else 0.0) == self.world

    def timed(self, fn, steps, warmup, n_marks):
        """Runs fn(events or None) warmup + steps times; returns (ms per step: max over ranks, per-leg ms on this rank)."""
        torch = self.torch
        for _ in range(warmup):
            fn(None)
        self.barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)] for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(self.stream)
        for i in range(steps):
            fn(ev[i])
        t1.record(self.stream)
        self.barrier()
        total = self.max_over_ranks(t0.elapsed_time(t1))
        legs = [float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(steps)])) for k in range(n_marks - 1)]
        return total / steps, legs

    def leg(self, ms, npx, alg_bytes, n_gpus=None):
        """One leg's report; `frac` divides by the measured HBM peak of all the GPUs that worked on it."""
        n = n_gpus or 1
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        return {"ms": ms, "mpx_s": npx / (ms * 1e-3) / 1e6, "gb_s": gbs, "algorithmic_bytes": int(alg_bytes),
                "frac_of_measured_hbm": gbs / (self.peak * n), "frac_of_nominal_8tbs": gbs / (8000.0 * n), "gpus": n}


LEGS = ["sqoa_encode", "sqoa_decode", "qoi_encode", "qoi_decode"]


def run_headline(env: Env, args):
    """cfg2, IMAGES_PER_STEP images per step through the batch entry points; e2e through sqoa_encode / sqoa_decode."""
    torch, sb = env.torch, env.sb
    from seqoia_b200 import synth

    w, h, ch = 3840, 2160, 3
    npx = w * h
    n = IMAGES_PER_STEP
    img = synth.cfg2()  # every rank: the same recipe (independent units; weak scaling)
    raw = npx * ch
    cap = (sb.max_stream_size(w, h, ch) + 63) // 64 * 64
    stride = (raw + 63) // 64 * 64
    ctx, sptr = env.ctx, env.sptr
    d_px = torch.empty(n * stride, dtype=torch.uint8, device=env.dev)
    src = torch.from_numpy(img.reshape(-1)).to(env.dev)
    for i in range(n):
        d_px[i * stride: i * stride + raw] = src
    d_stream = {q: torch.empty(n * cap, dtype=torch.uint8, device=env.dev) for q in (0, 1)}
    d_back = torch.empty(n * stride, dtype=torch.uint8, device=env.dev)
    d_len = {q: torch.zeros(n, dtype=torch.int32, device=env.dev) for q in (0, 1)}
    d_status = torch.zeros(n, dtype=torch.int32, device=env.dev)
    enc_plan = {q: ctx.plan([sb.Item(i * stride, i * cap, w, h, 0, ch, 0, q, 0) for i in range(n)]) for q in (0, 1)}
    for q in (0, 1):
        ctx.encode_batch(enc_plan[q], d_px, d_stream[q], d_len[q], sptr)
    torch.cuda.synchronize()
    slen = {q: int(d_len[q][0].item()) for q in (0, 1)}
    dec_plan = {q: ctx.plan([sb.Item(i * cap, i * stride, w, h, slen[q], ch, 0, q, ch) for i in range(n)], decode_=True)
                for q in (0, 1)}
    alg = {k: n * (raw + slen[0 if k.startswith("sqoa") else 1]) for k in LEGS}

    def step(ev):
        for k, (q, dec) in enumerate(((0, False), (0, True), (1, False), (1, True))):
            if ev:
                ev[k].record(env.stream)
            if dec:
                ctx.decode_batch(dec_plan[q], d_stream[q], d_back, d_status, sptr)
            else:
                ctx.encode_batch(enc_plan[q], d_px, d_stream[q], d_len[q], sptr)
        if ev:
            ev[4].record(env.stream)

    sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ms_per_step, leg_ms = env.timed(step, args.steps, max(3, args.warmup), 5)
    launches = (ctx.launches - launches0) * args.steps // (args.steps + max(3, args.warmup))
    clocks = sampler.stop() if env.rank == 0 else None
    # parity on the timed buffers (not timed): decoded pixels == input, stream digests == the reference's
    ok = bool(torch.equal(d_back.view(n, stride)[:, :raw], d_px.view(n, stride)[:, :raw])) and int(d_status.abs().sum().item()) == 0
    try:
        with open(os.path.join(ROOT, "tests", "golden", "digests.json")) as f:
            dig = json.load(f)["digests"]
        for q in (0, 1):
            want = dig[f"cfg2_3840x2160_rgb_q{q}"]
            got = bytes(d_stream[q][(n - 1) * cap: (n - 1) * cap + slen[q]].cpu().numpy())
            ok = ok and slen[q] == want["stream_len"] and hashlib.sha256(got).hexdigest() == want["stream_sha256"]
    except FileNotFoundError:
        pass

    # ---- e2e: the reference's own entry points on PAGEABLE host buffers (a C caller's malloc memory) ----
    import ctypes as C

    L = sb.lib()
    host_px = np.ascontiguousarray(img.reshape(-1)).copy()  # numpy: plain malloc memory
    h2d = d2h = 0
    e2e_secs = []
    e2e_ok = True
    e2e_images = 0
    # as many calling threads as the library keeps host contexts (SQOA_B200_HOST_CONTEXTS, default 2): one call's upload
    # overlaps another call's download, the way sqoabench's totals call the reference from several cores at once
    n_callers = max(1, min(n, sb.host_contexts()))

    def caller(count, check, acc):
        up = down = 0
        good = True
        for i in range(count):
            for q in (0, 1):
                d = sb.Desc(w, h, ch, 0, q)
                ln = C.c_int(0)
                sp = L.sqoa_encode(host_px.ctypes.data, C.byref(d), C.byref(ln))
                d2 = sb.Desc()
                pp = L.sqoa_decode(sp, ln.value, C.byref(d2), 0)
                up += host_px.size + ln.value
                down += ln.value + raw
                if not sp or not pp:
                    good = False
                elif check and i == 0:  # parity of the e2e path, outside the timed iterations
                    back = np.frombuffer(C.string_at(pp, raw), dtype=np.uint8)
                    good = good and bool(np.array_equal(back, host_px)) and ln.value == slen[q]
                L._free(sp)
                L._free(pp)
        acc.append((up, down, good))

    for it in range(1 + max(2, min(args.steps, 3))):
        acc = []
        if it == 0:  # warm-up: one image per caller, checked
            shares = [1] * n_callers
        else:
            shares = [n // n_callers + (1 if k < n % n_callers else 0) for k in range(n_callers)]
        threads = [threading.Thread(target=caller, args=(shares[k], it == 0, acc)) for k in range(n_callers)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        e2e_ok = e2e_ok and len(acc) == n_callers and all(a[2] for a in acc)
        if it > 0:
            h2d, d2h = sum(a[0] for a in acc), sum(a[1] for a in acc)
            e2e_secs.append(dt)
            e2e_images += n
    ok = ok and e2e_ok
    e2e_s = env.max_over_ranks(float(np.mean(e2e_secs)))
    ok = env.all_ok(ok)
    if env.rank != 0:
        return None
    world = env.world
    value = world * 4 * n * npx / (ms_per_step * 1e-3) / 1e6
    legs = {k: env.leg(leg_ms[i], n * npx, alg[k]) for i, k in enumerate(LEGS)}
    dom = max(LEGS, key=lambda k: legs[k]["ms"])
    traffic, traffic_src = traffic_of(dom)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": dict(CONFIG),
        "setup": {"images_per_step": n, "launches_per_step": 4,
                  "l2": f"{n} distinct images per step ({n * raw / 1e6:.0f} MB of pixels, > 3x L2): no leg finds its input in L2",
                  "parallelism": f"independent images, {world} GPU(s), no collective"},
        "legs": legs,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": legs[dom]["gb_s"], "peak": env.peak,
                     "unit": "GB/s", "frac": legs[dom]["gb_s"] / env.peak,
                     "traffic": traffic * n if traffic else None,
                     "traffic_note": f"dram read+write bytes of one launch ({n} images), ncu capture in profiles/{traffic_src}" if traffic else None,
                     "peak_kind": f"{env.peak_kind} copy bandwidth (MEASURED_PEAKS.json)",
                     "per_leg_frac": {k: legs[k]["frac_of_measured_hbm"] for k in LEGS}},
        "e2e": {"value": world * 4 * n * npx / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "iterations": e2e_images,
                "api": "sqoa_encode + sqoa_decode (seqoia.h:363,374) on pageable host buffers, both formats, one call per image, "
                       f"{n_callers} calling thread(s)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_spot_check": ok,
    }


def headline_cpu_baseline(line):
    try:
        from seqoia_b200 import synth

        codec = cpu_codec()
        img = synth.cfg2()
        w, h, ch = 3840, 2160, 3
        npx = w * h
        t1 = cpu_uniform_times(codec, img, w, h, ch, 1, 1)
        cores = os.cpu_count() or 1
        tn = cpu_uniform_times(codec, img, w, h, ch, cores, cores)
        line["cpu_baseline"] = {
            "value": 4 * npx / sum(t1.values()) / 1e6, "unit": UNIT, "cores": 1, "kind": codec.kind,
            "sample": "one cfg2 image, 4 legs, single thread",
            "legs_mpx_s": {k: npx / v / 1e6 for k, v in t1.items()},
            "all_cores": {"value": 4 * cores * npx / sum(tn.values()) / 1e6, "cores": cores,
                          "sample": f"{cores} copies, one image per core"},
        }
    except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}


# ---------------------------------------------------------------------------------------
# the other BASELINE.json configs (reported under "configs" of the same line)
# ---------------------------------------------------------------------------------------
def run_cfg1(env: Env, args):
    """configs[0]: one 1920x1080 RGBA image, SQOA round trip (sqoabench.c:434-455): latency, device and host API."""
    torch, sb = env.torch, env.sb
    import ctypes as C

    from seqoia_b200 import synth

    if env.rank != 0:
        return None
    w, h, ch = 1920, 1080, 4
    img = synth.cfg1()
    raw = w * h * ch
    ctx, sptr = env.ctx, env.sptr
    cap = sb.max_stream_size(w, h, ch)
    reps = 8  # rotate over buffers so that nothing is found in L2 (8 x (8.3 + 4.9 + 8.3 MB) > L2)
    d_px = [torch.from_numpy(img.reshape(-1)).to(env.dev) for _ in range(reps)]
    d_s = [torch.empty(cap + 64, dtype=torch.uint8, device=env.dev) for _ in range(reps)]
    d_o = [torch.empty(raw + 64, dtype=torch.uint8, device=env.dev) for _ in range(reps)]
    d_n = torch.zeros(4, dtype=torch.int32, device=env.dev)
    d_st = torch.zeros(4, dtype=torch.int32, device=env.dev)
    desc = sb.Desc(w, h, ch, 0, 0)
    for r in range(reps):
        ctx.encode_device(d_px[r], desc, d_s[r], cap, d_n, sptr)
    torch.cuda.synchronize()
    n = int(d_n[0].item())
    rc, dd, nb = sb.probe(bytes(d_s[0][:15].cpu().numpy()), n, 0)
    it = [0]

    def step(ev):
        r = it[0] % reps
        it[0] += 1
        if ev:
            ev[0].record(env.stream)
        ctx.encode_device(d_px[r], desc, d_s[r], cap, d_n, sptr)
        if ev:
            ev[1].record(env.stream)
        ctx.decode_device(d_s[(r + 1) % reps], n, dd, 0, d_o[r], raw, d_st, sptr)
        if ev:
            ev[2].record(env.stream)

    torch_ms, leg_ms = Env.timed(env, step, 20, 3, 3) if env.world == 1 else (None, None)
    if leg_ms is None:  # (under torchrun only rank 0 runs this: no collective inside)
        for _ in range(3):
            step(None)
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
        for e in evs:
            step(e)
        torch.cuda.synchronize()
        leg_ms = [float(np.mean([e[k].elapsed_time(e[k + 1]) for e in evs])) for k in range(2)]
    ok = bool(torch.equal(d_o[0][:raw], d_px[0]))
    dig = {}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "digests.json")) as f:
            dig = json.load(f)["digests"]["cfg1_1920x1080_rgba_q0"]
        ok = ok and n == dig["stream_len"] and hashlib.sha256(bytes(d_s[0][:n].cpu().numpy())).hexdigest() == dig["stream_sha256"]
    except Exception:
        pass
    # host API latency (pageable buffers)
    L = sb.lib()
    host_px = np.ascontiguousarray(img.reshape(-1)).copy()
    lat = []
    for i in range(12):
        d = sb.Desc(w, h, ch, 0, 0)
        ln = C.c_int(0)
        t0 = time.perf_counter()
        sp = L.sqoa_encode(host_px.ctypes.data, C.byref(d), C.byref(ln))
        t1 = time.perf_counter()
        d2 = sb.Desc()
        pp = L.sqoa_decode(sp, ln.value, C.byref(d2), 0)
        t2 = time.perf_counter()
        L._free(sp)
        L._free(pp)
        if i >= 2:
            lat.append((t1 - t0, t2 - t1))
    res = {"workload": "cfg1: one 1920x1080 RGBA image, SQOA encode + decode round trip",
           "legs": {"sqoa_encode": env.leg(leg_ms[0], w * h, raw + n), "sqoa_decode": env.leg(leg_ms[1], w * h, raw + n)},
           "latency_us": {"device_encode": leg_ms[0] * 1e3, "device_decode": leg_ms[1] * 1e3,
                          "host_api_encode": float(np.median([a for a, _ in lat])) * 1e6,
                          "host_api_decode": float(np.median([b for _, b in lat])) * 1e6},
           "e2e": {"value": 2 * w * h / float(np.median([a + b for a, b in lat])) / 1e6, "unit": UNIT,
                   "api": "sqoa_encode + sqoa_decode on pageable host buffers"},
           "parity": ok, "parity_against": "reference digest (tests/golden/digests.json) + round trip"}
    try:
        codec = cpu_codec()
        t = cpu_uniform_times(codec, img, w, h, ch, 1, 1)
        res["cpu_baseline"] = {"kind": codec.kind, "cores": 1, "sample": "the cfg1 image, single thread",
                               "sqoa_encode_mpx_s": w * h / t["sqoa_encode"] / 1e6, "sqoa_decode_mpx_s": w * h / t["sqoa_decode"] / 1e6,
                               "round_trip_ms": (t["sqoa_encode"] + t["sqoa_decode"]) * 1e3}
    except Exception as e:
        res["cpu_baseline"] = {"kind": "unavailable", "sample": str(e)}
    return res


def _concat_sha(d_arena, offs, lens, chunk=4096):
    """SHA-256 of the streams at arena[offs[i] : offs[i] + lens[i]] concatenated in order (device arena)."""
    hsh = hashlib.sha256()
    host = d_arena.cpu().numpy()
    for o, ln in zip(offs, lens):
        hsh.update(host[o:o + ln].data)
    return hsh


def run_cfg3(env: Env, args):
    """configs[2]: 100k 64x64 RGBA icons sharded by image index (no collective): encode + decode, both formats."""
    torch, sb = env.torch, env.sb
    from seqoia_b200 import dist as sdist
    from seqoia_b200 import synth

    n_total = args.images
    lo, hi = sdist.shard_range(n_total, env.world, env.rank)
    n = hi - lo
    icons = synth.cfg3(n, first=lo)
    px_bytes = 64 * 64 * 4
    cap = (sb.max_stream_size(64, 64, 4) + 63) // 64 * 64
    ctx, sptr = env.ctx, env.sptr
    d_px = torch.from_numpy(icons.reshape(-1)).to(env.dev)
    d_out = {q: torch.empty(n * cap, dtype=torch.uint8, device=env.dev) for q in (0, 1)}
    d_len = {q: torch.zeros(n, dtype=torch.int32, device=env.dev) for q in (0, 1)}
    d_back = torch.empty(n * px_bytes, dtype=torch.uint8, device=env.dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=env.dev)
    enc_plan = {q: ctx.plan([sb.Item(i * px_bytes, i * cap, 64, 64, 0, 4, 0, q, 0) for i in range(n)]) for q in (0, 1)}
    for q in (0, 1):
        ctx.encode_batch(enc_plan[q], d_px, d_out[q], d_len[q], sptr)
    torch.cuda.synchronize()
    lens = {q: d_len[q].cpu().numpy().astype(np.int64) for q in (0, 1)}
    dec_plan = {q: ctx.plan([sb.Item(i * cap, i * px_bytes, 64, 64, int(lens[q][i]), 4, 0, q, 4) for i in range(n)],
                            decode_=True) for q in (0, 1)}

    def step(ev):
        for k, (q, dec) in enumerate(((0, False), (0, True), (1, False), (1, True))):
            if ev:
                ev[k].record(env.stream)
            if dec:
                ctx.decode_batch(dec_plan[q], d_out[q], d_back, d_status, sptr)
            else:
                ctx.encode_batch(enc_plan[q], d_px, d_out[q], d_len[q], sptr)
        if ev:
            ev[4].record(env.stream)

    ms, leg_ms = env.timed(step, args.steps, 3, 5)
    ok = bool(torch.equal(d_back, d_px)) and int(d_status.abs().sum().item()) == 0
    # parity against the reference: SHA-256 of all streams in image order (ranks hash their own, rank 0 chains them)
    parity_against = "round trip"
    dig = golden_full()
    if n_total == 100_000 and f"cfg3_icons_0_{n_total - 1}_q0" in dig:
        parity_against = "reference digest of all 100,000 streams (tests/golden/digests_full.json) + round trip"
        for q in (0, 1):
            want = dig[f"cfg3_icons_0_{n_total - 1}_q{q}"]
            total_len = env.sum_over_ranks(float(lens[q].sum()))
            if env.world == 1:
                hsh = _concat_sha(d_out[q], [i * cap for i in range(n)], lens[q])
                ok = ok and hsh.hexdigest() == want["stream_sha256"]
            else:
                # gather the compacted streams on rank 0 (one host copy of this rank's arena, then slices)
                host = d_out[q].cpu().numpy().reshape(n, cap)
                mine = torch.from_numpy(np.concatenate([host[i, :lens[q][i]] for i in range(n)])).to(env.dev)
                del host
                sizes = [torch.zeros(1, dtype=torch.int64, device=env.dev) for _ in range(env.world)]
                env.dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device=env.dev))
                pad = max(int(s.item()) for s in sizes)
                buf = torch.zeros(pad, dtype=torch.uint8, device=env.dev)
                buf[: mine.numel()] = mine
                parts = [torch.empty(pad, dtype=torch.uint8, device=env.dev) for _ in range(env.world)] if env.rank == 0 else None
                env.dist.gather(buf, parts, dst=0)
                if env.rank == 0:
                    hsh = hashlib.sha256()
                    for r in range(env.world):
                        hsh.update(parts[r][: int(sizes[r].item())].cpu().numpy().data)
                    ok = ok and hsh.hexdigest() == want["stream_sha256"]
                del parts, buf, mine
            ok = ok and int(total_len) == want["stream_len"]
    ok = env.all_ok(ok)
    bytes_all = {q: env.sum_over_ranks(float(n * px_bytes + lens[q].sum())) for q in (0, 1)}
    leg_max = [env.max_over_ranks(v) for v in leg_ms]
    if os.environ.get("SQOA_BENCH_DEBUG"):
        print(f"[bench] cfg3 rank {env.rank}: legs {[round(v, 3) for v in leg_ms]} ms, stream bytes {[int(lens[q].sum()) for q in (0, 1)]}",
              file=sys.stderr, flush=True)
    if env.rank != 0:
        return None
    npx_all = n_total * 4096
    res = {"workload": f"cfg3: {n_total} 64x64 RGBA icons sharded by image index over {env.world} GPU(s), 4 legs per step",
           "scaling": "strong", "ms_per_step": ms, "value": 4 * npx_all / (ms * 1e-3) / 1e6, "unit": UNIT,
           "legs": {k: env.leg(leg_max[i], npx_all, bytes_all[0 if k.startswith("sqoa") else 1], env.world) for i, k in enumerate(LEGS)},
           "parity": ok, "parity_against": parity_against, "images_per_gpu": n}
    if env.world == 1:
        try:
            codec = cpu_codec()
            cores = os.cpu_count() or 1
            n1, nn = min(n, 2048), min(n, 16384)
            sample = icons[:n1]
            t1 = cpu_uniform_times(codec, sample, 64, 64, 4, n1, 1)
            big = icons[:nn]
            tn = cpu_uniform_times(codec, big, 64, 64, 4, nn, cores)
            res["cpu_baseline"] = {"kind": codec.kind,
                                   "single_thread": {"cores": 1, "sample": f"icons 0..{n1 - 1}", "mpx_s": 4 * n1 * 4096 / sum(t1.values()) / 1e6,
                                                     "legs_mpx_s": {k: n1 * 4096 / v / 1e6 for k, v in t1.items()}},
                                   "all_cores": {"cores": cores, "sample": f"icons 0..{nn - 1}, one image per core at a time",
                                                 "mpx_s": 4 * nn * 4096 / sum(tn.values()) / 1e6,
                                                 "legs_mpx_s": {k: nn * 4096 / v / 1e6 for k, v in tn.items()}}}
        except Exception as e:
            res["cpu_baseline"] = {"kind": "unavailable", "sample": str(e)}
    return res


def run_cfg5(env: Env, args):
    """configs[4]: mixed-size corpus mirroring the qoi test-suite mix; SQOA<->QOI transcode on the device.  Images are
    dealt to the GPUs largest first, each to the rank with the fewest pixels so far; no collective."""
    torch, sb = env.torch, env.sb
    from seqoia_b200 import synth

    shapes = synth.cfg5_shapes(args.scale)
    if env.world == 1:
        mine = shapes
    else:
        # largest image first, each to the rank with the fewest pixels so far: every rank gets the same pixels AND the
        # same mix of icons, screenshots and photos (contiguous index ranges gave one rank all the icons: 5.6 x at 8 GPUs)
        from seqoia_b200 import dist as sdist

        owner = sdist.deal_by_weight([w * h for _k, w, h, _c, _s in shapes], env.world)
        mine = [shapes[i] for i in range(len(shapes)) if owner[i] == env.rank]
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
    ctx, sptr = env.ctx, env.sptr
    d_px = torch.from_numpy(host).to(env.dev)
    d_st = {q: torch.zeros(st_total, dtype=torch.uint8, device=env.dev) for q in (0, 1)}   # direct encodes (sources)
    d_tr = {q: torch.zeros(st_total, dtype=torch.uint8, device=env.dev) for q in (0, 1)}   # transcoded streams
    d_len = {q: torch.zeros(n, dtype=torch.int32, device=env.dev) for q in (0, 1)}
    d_len_tr = {q: torch.zeros(n, dtype=torch.int32, device=env.dev) for q in (0, 1)}
    d_status = torch.zeros(n, dtype=torch.int32, device=env.dev)
    enc_plan = {q: ctx.plan([sb.Item(px_off[i], st_off[i], w, h, 0, c, 0, q, 0)
                             for i, (_k, w, h, c, _s) in enumerate(mine)]) for q in (0, 1)}
    for q in (0, 1):
        ctx.encode_batch(enc_plan[q], d_px, d_st[q], d_len[q], sptr)
    torch.cuda.synchronize()
    lens = {q: d_len[q].cpu().numpy().astype(np.int64) for q in (0, 1)}
    tplan = {}
    for src, dst in ((0, 1), (1, 0)):
        tplan[(src, dst)] = ctx.transcode_plan(
            [sb.Item(st_off[i], st_off[i], w, h, int(lens[src][i]), c, 0, src, c) for i, (_k, w, h, c, _s) in enumerate(mine)], dst)

    def step(ev):
        for k, (src, dst) in enumerate(((0, 1), (1, 0))):
            if ev:
                ev[k].record(env.stream)
            ctx.transcode_batch(tplan[(src, dst)], d_st[src], d_tr[dst], d_len_tr[dst], d_status, sptr)
        if ev:
            ev[2].record(env.stream)

    ms, leg_ms = env.timed(step, args.steps, 3, 3)
    # a transcoded stream must equal the direct encoding of the original pixels, byte for byte; the direct encodings
    # are compared with the reference's digest of the whole corpus (streams concatenated in image order)
    ok = all(bool(torch.equal(d_tr[q], d_st[q])) and bool(torch.equal(d_len_tr[q], d_len[q])) for q in (0, 1))
    ok = ok and int(d_status.abs().sum().item()) == 0
    parity_against = "transcoded stream == direct encoding of the original pixels"
    dig = golden_full()
    key = f"cfg5_scale{args.scale}_q0"
    if key in dig and env.world == 1:
        parity_against += "; direct encodings == reference digest of the corpus (tests/golden/digests_full.json)"
        for q in (0, 1):
            want = dig[f"cfg5_scale{args.scale}_q{q}"]
            ok = ok and int(lens[q].sum()) == want["stream_len"] and _concat_sha(d_tr[q], st_off, lens[q]).hexdigest() == want["stream_sha256"]
    ok = env.all_ok(ok)
    npx = int(sum(w * h for _k, w, h, _c, _s in mine))
    npx_all = env.sum_over_ranks(float(npx))
    bytes_all = env.sum_over_ranks(float(lens[0].sum() + lens[1].sum()))
    leg_max = [env.max_over_ranks(v) for v in leg_ms]
    if os.environ.get("SQOA_BENCH_DEBUG"):
        print(f"[bench] cfg5 rank {env.rank}: legs {[round(v, 3) for v in leg_ms]} ms, {n} images, {npx / 1e6:.1f} Mpx", file=sys.stderr, flush=True)
    if env.rank != 0:
        return None
    res = {"workload": f"cfg5: {len(shapes)} images of the qoi-suite mix (scale {args.scale}), {npx_all / 1e6:.0f} Mpx, "
                       f"SQOA->QOI and QOI->SQOA on the device over {env.world} GPU(s)",
           "scaling": "strong", "ms_per_step": ms, "value": 2 * npx_all / (ms * 1e-3) / 1e6, "unit": UNIT,
           "bytes": "stream in + stream out per direction (the pixels stay in a buffer that is reused group by group)",
           "legs": {name: env.leg(leg_max[i], npx_all, bytes_all, env.world) for i, name in enumerate(("sqoa_to_qoi", "qoi_to_sqoa"))},
           "parity": ok, "parity_against": parity_against}
    if env.world == 1:
        try:
            codec = cpu_codec()
            cores = os.cpu_count() or 1
            pick = list(range(0, n, max(1, n // 160)))  # a bounded sample across the whole mix
            streams = {i: bytes(d_st[0][st_off[i]: st_off[i] + int(lens[0][i])].cpu().numpy()) for i in pick}

            def one(i):
                _k, w, h, c, _s = mine[i]
                px, _d = codec.decode(streams[i], 0)
                codec.encode(px, w, h, c, 0, 1)
                return w * h

            t0 = time.perf_counter()
            px1 = sum(one(i) for i in pick[:: 8])
            t_single = time.perf_counter() - t0
            t0 = time.perf_counter()
            with concurrent.futures.ThreadPoolExecutor(cores) as ex:
                pxn = sum(ex.map(one, pick))
            t_all = time.perf_counter() - t0
            res["cpu_baseline"] = {"kind": codec.kind, "what": "SQOA -> QOI (sqoa_decode + sqoa_encode), malloc/free inside",
                                   "single_thread": {"cores": 1, "sample": f"{len(pick[::8])} images", "mpx_s": px1 / t_single / 1e6},
                                   "all_cores": {"cores": cores, "sample": f"{len(pick)} images, one image per core at a time",
                                                 "mpx_s": pxn / t_all / 1e6}}
        except Exception as e:
            res["cpu_baseline"] = {"kind": "unavailable", "sample": str(e)}
    return res


def run_cfg4(env: Env, args):
    """configs[3]: one 20000x19999 RGBA image.  N = 1: whole image.  N > 1: scanline-sharded, only boundary summaries
    cross GPUs (one all-gather of 320 bytes per rank, folded on the device); stream segments stay sharded."""
    torch, sb = env.torch, env.sb
    from seqoia_b200 import dist as sdist
    from seqoia_b200 import synth

    w, h = args.width, args.height
    world, rank = env.world, env.rank
    y0, y1 = sdist.shard_rows(h, world, rank)
    mine = synth.cfg4_rows(y0, y1, w, h)
    n_px = (y1 - y0) * w
    ctx, sptr = env.ctx, env.sptr
    d_px = torch.from_numpy(mine.reshape(-1)).to(env.dev)
    del mine
    cap = n_px * 5 + 64
    d_seg = torch.empty(cap, dtype=torch.uint8, device=env.dev)
    d_len = torch.zeros(4, dtype=torch.int32, device=env.dev)
    dig = golden_full()
    res, ok_all, parity_notes = {}, True, []
    steps = max(3, min(args.steps, 5))
    for q, name in ((0, "sqoa"), (1, "qoi")):
        desc = sb.Desc(w, h, 4, 0, q)
        enc = sdist.ShardedEncoder(ctx, env.dev, group=None) if world > 1 else None

        def encode(ev):
            if ev:
                ev[0].record(env.stream)
            if world > 1:
                enc.encode(d_px, n_px, desc, d_seg, cap, d_len, sptr)
            else:
                ctx.encode_device(d_px, desc, d_seg, cap, d_len, sptr)
            if ev:
                ev[1].record(env.stream)

        ms, _legs = env.timed(encode, steps, 3, 2)
        seg_len = int(d_len[0].item())
        stream_len = int(env.sum_over_ranks(float(seg_len)))
        res[f"{name}_encode"] = env.leg(ms, w * h, w * h * 4 + stream_len, world)
        res[f"{name}_encode"]["stream_bytes"] = stream_len
        # parity: SHA-256 of the whole stream (segments concatenated in rank order) against the reference's digest
        key = f"cfg4_{w}x{h}_rgba_q{q}"
        if world > 1:
            lens_all = [torch.zeros(1, dtype=torch.int64, device=env.dev) for _ in range(world)]
            env.dist.all_gather(lens_all, torch.tensor([seg_len], dtype=torch.int64, device=env.dev))
            lens_all = [int(x.item()) for x in lens_all]
            pad = max(lens_all)
            mine_seg = torch.zeros(pad, dtype=torch.uint8, device=env.dev)
            mine_seg[:seg_len] = d_seg[:seg_len]
            parts = [torch.empty(pad, dtype=torch.uint8, device=env.dev) for _ in range(world)]
            env.dist.all_gather(parts, mine_seg)
            full = torch.cat([parts[r][: lens_all[r]] for r in range(world)] + [torch.zeros(64, dtype=torch.uint8, device=env.dev)])
            del parts, mine_seg
        else:
            full = d_seg
        if key in dig:
            if rank == 0:
                hsh = hashlib.sha256()
                for a in range(0, stream_len, 256 << 20):
                    hsh.update(full[a:min(stream_len, a + (256 << 20))].cpu().numpy().data)
                good = stream_len == dig[key]["stream_len"] and hsh.hexdigest() == dig[key]["stream_sha256"]
                ok_all = ok_all and good
                parity_notes.append(f"{name} stream == reference digest: {good}")
        else:
            parity_notes.append(f"{name}: no reference digest for {w}x{h}")
        # decode
        if world == 1:
            rc, dd, nbytes = sb.probe(bytes(d_seg[:15].cpu().numpy()), seg_len, 0)
            d_back = torch.empty(nbytes + 64, dtype=torch.uint8, device=env.dev)
            d_st = torch.zeros(4, dtype=torch.int32, device=env.dev)

            def decode(ev):
                if ev:
                    ev[0].record(env.stream)
                ctx.decode_device(d_seg, seg_len, dd, 0, d_back, nbytes, d_st, sptr)
                if ev:
                    ev[1].record(env.stream)

            ms, _legs = env.timed(decode, steps, 3, 2)
            good = bool(torch.equal(d_back[:nbytes], d_px)) and int(d_st[0].item()) == 0
            ok_all = ok_all and good
            res[f"{name}_decode"] = env.leg(ms, w * h, w * h * 4 + stream_len, 1)
            res[f"{name}_decode"]["round_trip_ok"] = good
            del d_back
        else:
            # stream-sharded decode: every rank decodes one byte range of the stream (cut on decoder tile boundaries,
            # sqoa_b200_decode_sharded_device).  SQOA: three passes per rank, only 8-word summaries cross GPUs; QOI:
            # the ranges one after the other, the 544-byte decoder state handed from rank to rank
            total = stream_len
            hdr = 15 if q == 0 else 14
            cuts = sdist.stream_cuts(total - hdr - 8, world)
            b0, b1 = cuts[rank], cuts[rank + 1]
            avail = min(total - (hdr + b0), b1 - b0 + 32)
            if q == 1 and rank < world - 1:
                avail = b1 - b0 + 64  # (what follows the stream in `full` is zero padding)
            d_body = full[hdr + b0: hdr + b0 + avail + 64].clone()  # (its own allocation: the QOI path wants 16-byte alignment)
            dec = sdist.ShardedDecoder(ctx, env.dev, group=None)
            out = {}

            out["r"] = dec.decode(d_body, avail, b1 - b0, desc, 0, rank, world, sptr)  # sizes the pixel buffer

            def decode(ev):  # one library call, stream-ordered, nothing read back
                if ev:
                    ev[0].record(env.stream)
                dec.launch(d_body, avail, b1 - b0, desc, 0, sptr)
                if ev:
                    ev[1].record(env.stream)

            ms, _legs = env.timed(decode, steps, 3, 2)
            d_out, first_px, n_mine = out["r"]
            if int(dec.d_status.item()) != 0 or [int(v) for v in dec.d_info.tolist()] != [first_px, n_mine]:
                ok_all = False
            ya, yb = first_px // w, min(h, (first_px + n_mine + w - 1) // w)
            ref = synth.cfg4_rows(ya, yb, w, h).reshape(-1)[(first_px - ya * w) * 4: (first_px - ya * w + n_mine) * 4]
            good = bool(torch.equal(d_out[: n_mine * 4].cpu(), torch.from_numpy(ref.copy())))
            good = env.all_ok(good) and int(env.sum_over_ranks(float(n_mine))) == w * h
            ok_all = ok_all and good
            res[f"{name}_decode"] = env.leg(ms, w * h, w * h * 4 + total, world)
            res[f"{name}_decode"].update({"round_trip_ok": good, "passes": dec.describe(q == 1)})
            del d_body, dec, d_out
        del full
    ok_all = env.all_ok(ok_all)
    if rank != 0:
        return None
    out = {"workload": f"cfg4: one {w}x{h} RGBA image, " + ("whole image on one GPU" if world == 1 else
                       f"rows sharded over {world} GPUs; exchange = one all-gather of 320-byte shard summaries, folded on the device"),
           "scaling": "strong", "legs": res, "parity": ok_all, "parity_against": "; ".join(parity_notes) + "; decode: pixels == input"}
    if world == 1:
        try:
            codec = cpu_codec()
            rows = min(h, 2000)  # a bounded sample: the first 2000 scanlines as an image of their own
            band = synth.cfg4_rows(0, rows, w, h)
            t = cpu_uniform_times(codec, band, w, rows, 4, 1, 1)
            out["cpu_baseline"] = {"kind": codec.kind, "cores": 1,
                                   "sample": f"the first {rows} scanlines ({w * rows / 1e6:.0f} Mpx) as one image, single thread "
                                             "(a single image is inherently single-threaded in the reference)",
                                   "legs_mpx_s": {k: w * rows / v / 1e6 for k, v in t.items()}}
        except Exception as e:
            out["cpu_baseline"] = {"kind": "unavailable", "sample": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--only", default="", help="run one workload only: cfg2 (headline) | cfg1 | cfg3 | cfg4 | cfg5")
    ap.add_argument("--workload", default="", help="alias of --only")
    ap.add_argument("--skip-configs", action="store_true", help="headline only")
    ap.add_argument("--scale", type=float, default=1.0, help="cfg5: fraction of the 2,851-image corpus")
    ap.add_argument("--images", type=int, default=100_000, help="cfg3: images in the batch")
    ap.add_argument("--width", type=int, default=20000, help="cfg4")
    ap.add_argument("--height", type=int, default=19999, help="cfg4")
    args = ap.parse_args()
    only = args.only or args.workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    env = Env(rank, world, local_rank)
    runners = {"cfg1": run_cfg1, "cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5}
    line = None
    if only in ("", "cfg2"):
        line = run_headline(env, args)
        if rank == 0 and world == 1:
            headline_cpu_baseline(line)
    if only in runners:
        r = runners[only](env, args)
        if rank == 0:
            print(json.dumps(r), flush=True)
        return
    if only == "" and not args.skip_configs:
        configs = {}
        for name in ("cfg1", "cfg3", "cfg4", "cfg5"):
            t0 = time.time()
            if rank == 0:
                print(f"[bench] {name} ...", file=sys.stderr, flush=True)
            try:
                r = runners[name](env, args)
            except Exception as e:  # a failing side workload must not lose the headline (reported, not hidden)
                r = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise
            if rank == 0 and r is not None:
                r["wall_s"] = round(time.time() - t0, 1)
                configs[name] = r
            env.torch.cuda.empty_cache()
        if rank == 0:
            line["configs"] = configs
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
