#!/usr/bin/env python
"""bench.py — Delta-Rice encode+decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input: ENCODE the batch
(raw int16 -> Delta-Rice chunk streams), [N>1: all-gather the per-shard byte counts + scan],
then DECODE the streams back (-> raw int16); the all-gather runs on a side stream next to the
decode kernels and is joined before the step ends.  Workload at every N (weak scaling, one
process per GPU): BASELINE.json configs[1] "C2" per GPU — 153 391 Nab-like waveforms of
3500 samples (1.074 GB raw), RiceParameter M=4, chunks of 2000 waveforms.

`value` = raw int16 bytes pushed through the codec per second, counting both directions:
    value = N_gpus * 2 * raw_bytes_per_gpu / t_step          (GB/s, 1e9)
i.e. the harmonic combination of encode GB/s and decode GB/s, both also printed
(`encode_gbs`, `decode_gbs`).  Inputs are resident in HBM when the timed region starts and
are larger than L2 (1.07 GB raw + 0.29 GB stream vs 126 MB), so no flush is needed.
`e2e` = the same metric through the host-pointer C-ABI (drice_encode_batch_host /
drice_decode_batch_host: what H5Z_filter_deltarice calls) with pinned HOST buffers, every
step copying raw in + stream out and stream in + raw out.  Headline: two handles on two host
threads, step s encodes while the streams of step s-1 are decoded (both PCIe directions busy);
`e2e.sequential` = one handle, encode then decode.
`--impl reference` times the UNMODIFIED reference (oracle/_ref/libref_omp.so, its own
H5Z_filter_deltarice, OpenMP over all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encode+decode GB/s of raw int16 (bit-exact Delta-Rice, HDF5 filter 32025 stream)"
UNIT = "GB/s"
WORKLOADS = {
    # name: (n_waves, L, M, waves_per_chunk, generator)
    "c2": (153391, 3500, 4, 2000, "nab"),        # BASELINE.json configs[1] — the bench line
    "c3m8": (76696, 7000, 8, 2000, "nab"),       # one 1 GiB slice of configs[2] at M=8
    "c4": (76696, 7000, 8, 2000, "mix"),         # configs[3] input mix (decode stress)
    "c1": (100, 7000, 8, 20, "gauss10"),         # README case (tiny; parity config)
}


def workload_config(name, world):
    n_waves, L, M, wpc, gen = WORKLOADS[name]
    return {
        "workload": f"{name}: {n_waves} {gen} waveforms x {L} int16 per GPU, RiceParameter={M}, "
                    f"WaveformLength={L}, chunks of {wpc} waveforms",
        "raw_bytes_per_gpu": n_waves * L * 2, "n_waves": n_waves, "L": L, "M": M,
        "waves_per_chunk": wpc, "sharding": f"{world} rank(s), whole chunks per rank, weak",
        "cache": "inputs larger than L2 (no flush needed)",
    }


# ======================================================================================
# reference arm (CPU): numpy + ctypes only, no torch, no CUDA
# ======================================================================================
def _ref_sample(name, sample_chunks, seed=20251018):
    from deltarice_b200.synth import nab_like, gaussian_mix
    n_waves, L, M, wpc, gen = WORKLOADS[name]
    wpc = min(wpc, n_waves)
    nw = min(n_waves, sample_chunks * wpc)
    if gen == "nab":
        x = nab_like(nw, L, seed)
    elif gen == "mix":
        x = gaussian_mix(nw, L, seed=seed)
    else:
        x = np.random.default_rng(0).normal(0, 10, (nw, L)).astype(np.int16)
    chunks = [np.ascontiguousarray(x[i:i + wpc]).ravel() for i in range(0, nw, wpc)]
    return chunks, L, M


class _RefRunner:
    """Drives H5Z_filter_deltarice of a CPU library exactly as libhdf5 would: malloc'ed
    *buf, ownership handed over.  Input staging (malloc + memcpy) is outside the timing."""

    def __init__(self, kind):
        from oracle import oracle as O
        self.O = O
        self.kind = kind            # "reference" (oracle/_ref) or "port" (oracle restatement)
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]

    def _filter(self, lib, data: np.ndarray, cd, reverse):
        n = data.nbytes
        p = self.libc.malloc(n + 64)
        C.memmove(p, data.ctypes.data, n)
        C.memset(p + n, 0, 64)
        buf, bs = C.c_void_p(p), C.c_size_t(n)
        cdv = (C.c_uint * len(cd))(*cd)
        t0 = time.perf_counter()
        ret = lib.H5Z_filter_deltarice(0x100 if reverse else 0, len(cd), cdv, n, C.byref(bs), C.byref(buf))
        dt = time.perf_counter() - t0
        if ret in (0, C.c_size_t(-1).value):
            raise RuntimeError("reference filter failed")
        out = np.frombuffer(C.string_at(buf.value, ret), dtype=np.uint8)
        self.libc.free(buf)
        return out, dt

    def step(self, chunks, L, M, serial=False):
        """encode then decode every chunk; returns (t_enc, t_dec, compressed bytes)."""
        te = td = 0.0
        comp_bytes = 0
        if self.kind == "reference":
            lib = self.O.ref_lib("ser" if serial else "omp")
            streams = []
            for x in chunks:
                s, dt = self._filter(lib, x, (M, L), False)
                te += dt
                streams.append(s)
                comp_bytes += s.size
            for x, s in zip(chunks, streams):
                y, dt = self._filter(lib, s, (M, L), True)
                td += dt
                assert y.size == x.nbytes
            assert np.array_equal(y.view(np.int16), chunks[-1])
        else:
            for x in chunks:
                t0 = time.perf_counter()
                s = self.O.encode_chunk(x, M, L, mt=not serial)
                t1 = time.perf_counter()
                y = self.O.decode_chunk(s, M, L, mt=not serial)
                t2 = time.perf_counter()
                te += t1 - t0
                td += t2 - t1
                comp_bytes += s.nbytes
            assert np.array_equal(y, chunks[-1])
        return te, td, comp_bytes


def run_reference(args):
    """Times the reference CPU implementation; prints one JSON line (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as O
    serial = bool(args.ref_serial)
    kind = "reference" if O.ref_available("ser" if serial else "omp") else "port"
    if kind == "port":
        O.lib()
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the all-core arm must not inherit it.  The
    # variable is set before the OpenMP runtime is loaded AND the team size is set through the
    # runtime's own API afterwards; `cores` is what omp_get_max_threads() then reports.
    want = 1 if serial else len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(want)
    chunks, L, M = _ref_sample(args.workload, args.ref_chunks)
    raw = sum(c.nbytes for c in chunks)
    rr = _RefRunner(kind)
    (O.ref_lib("ser" if serial else "omp") if kind == "reference" else O.lib())
    cores = want
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(C.c_int(want))
        gomp.omp_get_max_threads.restype = C.c_int
        cores = int(gomp.omp_get_max_threads())
    except OSError:
        pass
    if serial:
        cores = 1
    for _ in range(args.warmup):
        rr.step(chunks, L, M, serial)
    te = td = 0.0
    comp = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, comp = rr.step(chunks, L, M, serial)
        te += a
        td += b
    wall = time.perf_counter() - t0
    t = te + td
    val = 2 * raw * args.steps / t / 1e9
    cfg = workload_config(args.workload, 1)
    sample = (f"{len(chunks)} chunks ({raw / 1e6:.0f} MB raw) of the workload per step, one "
              f"H5Z_filter_deltarice call per chunk, encode then decode, {args.steps} steps; "
              f"timed around the filter calls (its own mallocs included, input staging excluded)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/u32",
        "data": "synthetic", "config": cfg,
        "encode_gbs": round(raw * args.steps / te / 1e9, 4), "decode_gbs": round(raw * args.steps / td / 1e9, 4),
        "ratio": round(comp / raw, 5),
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 2),
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_subprocess(workload, serial, chunks, steps, warmup):
    """Runs the reference arm in a fresh process (no torch/CUDA in it; its OpenMP runtime
    owns the host cores) and returns its parsed JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
           "--steps", str(steps), "--warmup", str(warmup), "--ref-chunks", str(chunks)]
    if serial:
        cmd.append("--ref-serial")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS")}
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("cpu baseline produced no JSON: " + out.stderr[-400:])


# ======================================================================================
# clocks sampler (NVML in a thread; the timed regions are tens of ms .. seconds)
# ======================================================================================
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, uuid):
        self.samples, self.reasons, self.ok = [], set(), False
        self.sm_max = None
        self._stop = threading.Event()
        self._live = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self.th = threading.Thread(target=self._run, daemon=True)
        if self.ok:
            self.th.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._live.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                        nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                        nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, nm in self.REASONS.items():
                        if r & bit and nm != "gpu_idle":
                            self.reasons.add(nm)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def live(self, on):
        (self._live.set if on else self._live.clear)()

    def result(self):
        self._stop.set()
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ======================================================================================
# our arm
# ======================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries ONE JSON line: everything else that libraries print to fd 1 (NCCL's version
    # banner, ...) goes to stderr; the line itself is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — deltarice_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import deltarice_b200 as d
    from deltarice_b200 import shard, synth

    n_waves, L, M, wpc, gen = WORKLOADS[args.workload]
    wpc = min(wpc, n_waves)
    seed = 20251018 + rank
    if gen == "nab":
        x = synth.nab_like_torch(n_waves, L, seed, dev).reshape(-1)
    elif gen == "mix":
        x = synth.gaussian_mix_torch(n_waves, L, seed=seed, device=dev).reshape(-1)
    else:
        x = torch.from_numpy(np.random.default_rng(0).normal(0, 10, (n_waves, L)).astype(np.int16)).to(dev).reshape(-1)
    raw_bytes = x.numel() * 2
    off = d.chunk_offsets(wpc * L, x.numel())
    nchunks = len(off) - 1

    codec = d.DeltaRice(local_rank)
    cap = codec.bound_bytes(off, L)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_boff = torch.zeros(nchunks + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(2, dtype=torch.int32, device=dev)
    y = torch.empty_like(x)

    # one checked pass: offsets for the decode calls, round trip, status
    codec.encode_device_async(x, off, M, L, out, d_boff, d_status)
    boff = d_boff.cpu().numpy().astype(np.uint64)
    assert int(d_status[0]) == 0, "encode status"
    comp_bytes = int(boff[-1])
    comp = out[:comp_bytes]
    codec.decode_device_async(comp, boff, off, M, L, y, d_status)
    torch.cuda.synchronize()
    assert int(d_status[0]) == 0 and torch.equal(x, y), "decode(encode(x)) != x"
    ratio = comp_bytes / raw_bytes

    # the path's one exchange (per-shard byte counts -> offsets in the concatenated stream) needs the
    # encoder's result but nothing of the decoder: it runs on a side stream next to the decode kernels
    # and is joined before the step ends
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    def exchange():
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            shard.gather_shard_offsets(d_boff[nchunks:nchunks + 1])

    def step():
        codec.encode_device_async(x, off, M, L, out, d_boff, d_status)
        if world > 1:
            exchange()
        codec.decode_device_async(comp, boff, off, M, L, y, d_status)
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    props = torch.cuda.get_device_properties(dev)
    uuid = "GPU-" + str(props.uuid) if not str(props.uuid).startswith("GPU-") else str(props.uuid)
    clocks = ClockSampler(uuid)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    codec.timing(True)
    codec.timing_read(reset=True)
    l0 = codec.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    clocks.live(True)
    barrier()
    ev[0].record()
    for s in range(args.steps):
        codec.encode_device_async(x, off, M, L, out, d_boff, d_status)
        ev[2 * s + 1].record()
        if world > 1:
            exchange()
        codec.decode_device_async(comp, boff, off, M, L, y, d_status)
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)
        ev[2 * s + 2].record()
    barrier()
    clocks.live(False)
    launches = codec.launches - l0
    t_total = ev[0].elapsed_time(ev[-1])                       # ms, device clock
    t_enc = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps))
    t_dec = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps))
    ktimes = codec.timing_read(reset=True)
    codec.timing(False)
    assert int(d_status[0]) == 0

    # ---- e2e: host buffers through the chunk scheduler (what the H5Z callback calls) ----
    # (a) sequential: encode the batch, then decode its streams (one handle, one direction of
    #     PCIe busy at a time);
    # (b) pipelined (the headline): two handles on two host threads - step s encodes the batch into
    #     stream buffer s%2 while the streams step s-1 produced are decoded, so raw-in and raw-out
    #     share the link's two directions.  Every step still encodes one batch and decodes one.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_raw = codec.pinned_empty(x.numel(), np.int16)
    h_raw[:] = x.cpu().numpy()
    h_comp = [codec.pinned_empty(cap, np.uint8) for _ in range(2)]
    h_back = codec.pinned_empty(x.numel(), np.int16)
    h_boff = [np.zeros(nchunks + 1, dtype=np.uint64) for _ in range(2)]
    for i in range(2):
        nb = codec.encode_host_into(h_raw, off, M, L, h_comp[i], h_boff[i])
        codec.decode_host_into(h_comp[i][:nb], h_boff[i], off, M, L, h_back)
    assert nb == comp_bytes and np.array_equal(h_back, h_raw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nb = codec.encode_host_into(h_raw, off, M, L, h_comp[0], h_boff[0])
        codec.decode_host_into(h_comp[0][:nb], h_boff[0], off, M, L, h_back)
    torch.cuda.synchronize()
    t_e2e_seq = (time.perf_counter() - t0) * 1e3 / e2e_steps   # ms per step (wall: host work is part of it)

    codec2 = d.DeltaRice(local_rank)
    h_back[:] = 0

    def _enc(i):
        codec.encode_host_into(h_raw, off, M, L, h_comp[i], h_boff[i])

    def _dec(i):
        codec2.decode_host_into(h_comp[i][:comp_bytes], h_boff[i], off, M, L, h_back)

    def _pipelined(nsteps):
        for s_ in range(nsteps):
            ta = threading.Thread(target=_enc, args=(s_ & 1,))
            tb = threading.Thread(target=_dec, args=((s_ + 1) & 1,))
            ta.start(); tb.start()
            ta.join(); tb.join()

    _pipelined(2)
    barrier()
    clocks.live(True)
    t0 = time.perf_counter()
    _pipelined(e2e_steps)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    assert np.array_equal(h_back, h_raw), "pipelined e2e: decode(encode(x)) != x"
    codec2.close()
    clocks.live(False)
    clk = clocks.result()

    # ---- max over ranks ----
    tt = torch.tensor([t_total, t_enc, t_dec, t_e2e, t_e2e_seq], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_enc, t_dec, t_e2e, t_e2e_seq = [float(v) for v in tt.cpu()]
    ms_step = t_total / args.steps
    value = world * 2 * raw_bytes / (ms_step * 1e6)
    e2e_val = world * 2 * raw_bytes / (t_e2e * 1e6)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg = raw_bytes * (1.0 + ratio)                         # read raw + write stream (encode) / the reverse (decode)
        kern = {}
        for name, (ms, cnt) in ktimes.items():
            if cnt:
                kern[name] = {"launches": cnt, "avg_ms": round(ms / cnt, 4)}
        dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches"]) if kern else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:  # noqa: BLE001
            pass
        roof = None
        if dom:
            alg_dom = 4.0 * (nchunks + n_waves) if dom == "locate_kernel" else alg
            ach = alg_dom / (kern[dom]["avg_ms"] * 1e6)
            roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(alg_dom),
                    "all_kernels": {k: dict(v, achieved_gbs=round((4.0 * (nchunks + n_waves) if k == "locate_kernel" else alg) / (v["avg_ms"] * 1e6), 1))
                                    for k, v in kern.items()}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                allc = cpu_baseline_subprocess(args.workload, False, args.ref_chunks, 3, 1)
                one = cpu_baseline_subprocess(args.workload, True, max(1, args.ref_chunks // 4), 1, 1)
                cpu = dict(allc["cpu_baseline"])
                cpu["encode_gbs"], cpu["decode_gbs"] = allc["encode_gbs"], allc["decode_gbs"]
                cpu["single_thread"] = {"value": one["value"], "encode_gbs": one["encode_gbs"],
                                        "decode_gbs": one["decode_gbs"], "cores": 1, "kind": one["cpu_baseline"]["kind"]}
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: " + repr(e)[:200]}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16/u32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "encode_gbs": round(world * raw_bytes * args.steps / (t_enc * 1e6), 1),
            "decode_gbs": round(world * raw_bytes * args.steps / (t_dec * 1e6), 1),
            "ratio": round(ratio, 5),
            "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": raw_bytes + comp_bytes,
                    "d2h_bytes_per_step": comp_bytes + raw_bytes, "steps": e2e_steps, "ms_per_step": round(t_e2e, 2),
                    "api": "drice_encode_batch_host + drice_decode_batch_host (pinned host buffers), two handles: "
                           "step s encodes while the streams of step s-1 are decoded (both PCIe directions busy)",
                    "sequential": {"value": round(world * 2 * raw_bytes / (t_e2e_seq * 1e6), 2), "ms_per_step": round(t_e2e_seq, 2),
                                   "api": "one handle: encode the batch, then decode its streams"}},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    codec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-chunks", type=int, default=8, help="chunks per step of the CPU reference sample")
    ap.add_argument("--ref-serial", action="store_true", help="reference arm single-threaded (serial build)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: launch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 500),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
