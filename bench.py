#!/usr/bin/env python
"""bench.py — Delta-Rice encode+decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|c1]

A "step" is one pass of the hot path over one batch of synthetic input: ENCODE the batch
(raw int16 -> Delta-Rice chunk streams), [N>1: all-gather the per-shard byte counts + scan],
then DECODE the streams back (-> raw int16); the all-gather runs on a side stream next to the
decode kernels and is joined before the step ends.

Workloads (BASELINE.json `configs`):
  c2  (default, the bench line; weak scaling) 153 391 Nab-like waveforms of 3500 samples per GPU
      (1.074 GB raw), RiceParameter 4, chunks of 2000 waveforms.
  c5  (strong scaling) ONE 64 GiB dataset, L=7000, M=8, 2454 chunks of 2000 waveforms generated from a
      per-chunk seed; rank r takes shard_chunk_range(nchunks, N, r).  After the timed steps the shards
      are gathered at the all-gathered offsets into one stream on rank 0 and checked: its header chain
      is walked on the host and sampled chunks are compared with the oracle's stream (`parity`).
  c3  RiceParameter sweep M = 1..64 at L=7000 on 8 GiB (1 GPU): one entry per M in `sweep`.
  c4  decode-only on a pre-compressed stream of mixed noise levels (sigma 1..1000, M=8, L=7000),
      `--c4-gib` GiB of compressed stream (default 32), decoded by ONE call.

`value` = raw int16 bytes pushed through the codec per second, counting both directions:
    value = 2 * raw_bytes_all_ranks / t_step          (GB/s, 1e9)
i.e. the harmonic combination of encode GB/s and decode GB/s, both also printed.  Inputs are
resident in HBM when the timed region starts and are far larger than L2 (no flush needed).
`e2e` = the same metric through the host-pointer C-ABI (drice_encode_batch_host /
drice_decode_batch_host) with pinned HOST buffers, every step copying raw in + stream out and
stream in + raw out; `e2e.h5z` = the same chunks through H5Z_filter_deltarice itself, one call per
chunk with malloc'ed buffers exactly as libhdf5 (and the reference arm) drives the filter.
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
GIB = 1 << 30
WORKLOADS = {
    # name: (n_waves, L, M, waves_per_chunk, generator, scaling)
    "c2": (153391, 3500, 4, 2000, "nab", "weak"),            # BASELINE.json configs[1] — the bench line
    "c3": (8 * GIB // 14000, 7000, 8, 2000, "nab", "weak"),  # configs[2]: 613 566 waves = 8 GiB, M swept
    "c4": (0, 7000, 8, 2000, "mix", "weak"),                 # configs[3]: size set by --c4-gib (compressed)
    "c5": (2454 * 2000, 7000, 8, 2000, "nab", "strong"),     # configs[4]: 64 GiB over all ranks
    "c1": (100, 7000, 8, 20, "gauss10", "weak"),             # README case (tiny; parity config)
}
C3_SWEEP = (1, 2, 4, 8, 16, 32, 64)
SEED = 20251018


def workload_config(name, world, n_waves=None):
    nw, L, M, wpc, gen, scaling = WORKLOADS[name]
    if n_waves is not None:
        nw = n_waves
    per = "in total over all ranks" if scaling == "strong" else "per GPU"
    return {
        "workload": f"{name}: {nw} {gen} waveforms x {L} int16 {per}, RiceParameter={M}, "
                    f"WaveformLength={L}, chunks of {wpc} waveforms",
        "raw_bytes": nw * L * 2, "n_waves": nw, "L": L, "M": M, "waves_per_chunk": wpc,
        "sharding": f"{world} rank(s), contiguous ranges of whole chunks per rank, {scaling}",
        "cache": "inputs larger than L2 (no flush needed)",
    }


# ======================================================================================
# H5Z callback driver (libhdf5's call pattern): used by the reference arm and by e2e.h5z
# ======================================================================================
class FilterRunner:
    """Drives an H5Z_filter_deltarice exactly as libhdf5 would: malloc'ed *buf, ownership handed
    over.  Input staging (malloc + memcpy) is outside the timing."""

    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]

    def call(self, lib, data: np.ndarray, cd, reverse):
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
            raise RuntimeError("H5Z_filter_deltarice failed")
        out = np.frombuffer(C.string_at(buf.value, ret), dtype=np.uint8)
        self.libc.free(buf)
        return out, dt

    def step(self, lib, chunks, L, M):
        """encode then decode every chunk through the callback; returns (t_enc, t_dec, compressed bytes)."""
        te = td = 0.0
        streams = []
        for x in chunks:
            s, dt = self.call(lib, x, (M, L), False)
            te += dt
            streams.append(s)
        y = None
        for x, s in zip(chunks, streams):
            y, dt = self.call(lib, s, (M, L), True)
            td += dt
            assert y.size == x.nbytes
        assert np.array_equal(y.view(np.int16), chunks[-1])
        return te, td, sum(s.size for s in streams)


def _ref_sample(name, sample_chunks, seed=SEED):
    """The chunks every CPU-timed leg works on: the first `sample_chunks` chunks of rank 0's input."""
    from deltarice_b200.synth import nab_like, gaussian_mix
    n_waves, L, M, wpc, gen, _ = WORKLOADS[name]
    if name == "c4":
        n_waves = sample_chunks * wpc
    wpc = min(wpc, n_waves)
    nw = min(n_waves, sample_chunks * wpc)
    if name == "c5":
        chunks = [nab_like(wpc, L, seed + c).ravel() for c in range(nw // wpc)]
        return chunks, L, M
    if gen == "nab":
        x = nab_like(nw, L, seed)
    elif gen == "mix":
        x = gaussian_mix(nw, L, seed=seed)
    else:
        x = np.random.default_rng(0).normal(0, 10, (nw, L)).astype(np.int16)
    chunks = [np.ascontiguousarray(x[i:i + wpc]).ravel() for i in range(0, nw, wpc)]
    return chunks, L, M


# ======================================================================================
# reference arm (CPU): numpy + ctypes only, no torch, no CUDA
# ======================================================================================
def run_reference(args):
    """Times the reference CPU implementation; prints one JSON line (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as O
    serial = bool(args.ref_serial)
    kind = "reference" if O.ref_available("ser" if serial else "omp") else "port"
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the all-core arm must not inherit it.  The
    # variable is set before the OpenMP runtime is loaded AND the team size is set through the
    # runtime's own API afterwards; `cores` is what omp_get_max_threads() then reports.
    want = 1 if serial else len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(want)
    chunks, L, M = _ref_sample(args.workload, args.ref_chunks)
    raw = sum(c.nbytes for c in chunks)
    lib = O.ref_lib("ser" if serial else "omp") if kind == "reference" else None
    if kind == "port":
        O.lib()
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
    fr = FilterRunner()

    def step():
        if kind == "reference":
            return fr.step(lib, chunks, L, M)
        te = td = 0.0
        comp = 0
        for x in chunks:
            t0 = time.perf_counter()
            s = O.encode_chunk(x, M, L, mt=not serial)
            t1 = time.perf_counter()
            y = O.decode_chunk(s, M, L, mt=not serial)
            t2 = time.perf_counter()
            te += t1 - t0
            td += t2 - t1
            comp += s.nbytes
        assert np.array_equal(y, chunks[-1])
        return te, td, comp

    for _ in range(args.warmup):
        step()
    te = td = 0.0
    comp = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, comp = step()
        te += a
        td += b
    wall = time.perf_counter() - t0
    t = te + td
    val = 2 * raw * args.steps / t / 1e9
    cfg = workload_config(args.workload, max(1, args.gpus))
    sample = (f"{len(chunks)} chunks ({raw / 1e6:.0f} MB raw) of the workload per step, one "
              f"H5Z_filter_deltarice call per chunk, encode then decode, {args.steps} steps; "
              f"timed around the filter calls (its own mallocs included, input staging excluded)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
        "higher_is_better": True, "scaling": WORKLOADS[args.workload][5], "vs_baseline": None, "dtype": "int16/u32",
        "data": "synthetic", "config": cfg,
        "encode_gbs": round(raw * args.steps / te / 1e9, 4), "decode_gbs": round(raw * args.steps / td / 1e9, 4),
        "ratio": round(comp / raw, 5),
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "omp_threads": cores},
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


def pin_rank_to_cores(rank_on_node, ranks_on_node):
    """Gives every rank of the node its own slice of the host cores it may run on (its Python
    threads, the library's copy threads and the first-touch placement of its pinned buffers then stay
    together).  Returns the cores kept."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if ranks_on_node <= 1 or len(cores) < 2 * ranks_on_node or os.environ.get("DRICE_NO_PIN"):
            return cores
        per = len(cores) // ranks_on_node
        mine = cores[rank_on_node * per:(rank_on_node + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return []


def host_copy_gbs(nbytes=1 << 28):
    """memcpy bandwidth of this rank's cores (bytes copied per second)."""
    a = np.ones(nbytes, dtype=np.uint8)
    b = np.empty_like(a)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        np.copyto(b, a)
        best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


def pcie_duplex_gbs(dev, barrier, nbytes=1 << 28, reps=3):
    """Pinned host <-> device copy bandwidth of this rank with BOTH directions busy, every rank of the node
    copying at the same time (the call is bracketed by barriers): the machine-level ceiling of the e2e legs,
    whose steps move raw + stream in each direction.  Returns (h2d GB/s, d2h GB/s) of this rank."""
    import torch
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = (0.0, 0.0)
    for _ in range(reps + 1):
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s1):
            e[0].record()
            d_in.copy_(h_in, non_blocking=True)
            e[1].record()
        with torch.cuda.stream(s2):
            e[2].record()
            h_out.copy_(d_out, non_blocking=True)
            e[3].record()
        torch.cuda.synchronize()
        cur = (nbytes / (e[0].elapsed_time(e[1]) * 1e6), nbytes / (e[2].elapsed_time(e[3]) * 1e6))
        if cur[0] + cur[1] > best[0] + best[1]:
            best = cur
    barrier()
    return best


# ======================================================================================
# our arm
# ======================================================================================
def make_input(name, rank, world, dev, args):
    """Returns (x, off, M, L, n_waves_all_ranks, first_chunk): this rank's samples on the device."""
    import torch
    import deltarice_b200 as d
    from deltarice_b200 import shard, synth
    n_waves, L, M, wpc, gen, scaling = WORKLOADS[name]
    first_chunk = 0
    if name == "c5":
        nchunks_all = n_waves // wpc
        if args.c5_chunks:
            nchunks_all = args.c5_chunks
        c0, c1 = shard.shard_chunk_range(nchunks_all, world, rank)
        x = torch.empty((c1 - c0) * wpc * L, dtype=torch.int16, device=dev)
        for c in range(c0, c1):                         # one generator state per GLOBAL chunk index
            x[(c - c0) * wpc * L:(c - c0 + 1) * wpc * L] = synth.nab_like_torch(wpc, L, SEED + c, dev).reshape(-1)
        off = d.chunk_offsets(wpc * L, x.numel())
        return x, off, M, L, nchunks_all * wpc, c0
    wpc = min(wpc, n_waves)
    seed = SEED + rank
    if gen == "nab":
        x = synth.nab_like_torch(n_waves, L, seed, dev).reshape(-1)
    elif gen == "mix":
        x = synth.gaussian_mix_torch(n_waves, L, seed=seed, device=dev).reshape(-1)
    else:
        x = torch.from_numpy(np.random.default_rng(0).normal(0, 10, (n_waves, L)).astype(np.int16)).to(dev).reshape(-1)
    off = d.chunk_offsets(wpc * L, x.numel())
    return x, off, M, L, n_waves * world, first_chunk


def run_c4(args, dev, codec):
    """decode-only: `--c4-gib` GiB of compressed stream (mixed noise), built slice by slice, decoded by one call."""
    import torch
    import deltarice_b200 as d
    from deltarice_b200 import synth
    from oracle import oracle as O
    _, L, M, wpc, _, _ = WORKLOADS["c4"]
    target = int(args.c4_gib * GIB)
    slice_waves = 36 * wpc                              # ~1 GB raw per slice
    cap = target + 2 * slice_waves * L * 2
    comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(2, dtype=torch.int32, device=dev)
    boffs, pos, nsl = [np.zeros(1, np.uint64)], 0, 0
    checked = 0
    last_x = None
    while pos < target:
        x = synth.gaussian_mix_torch(slice_waves, L, seed=SEED + nsl, device=dev).reshape(-1)
        off = d.chunk_offsets(wpc * L, x.numel())
        d_boff = torch.zeros(len(off), dtype=torch.int64, device=dev)
        codec.encode_device_async(x, off, M, L, comp[pos:], d_boff, d_status)
        b = d_boff.cpu().numpy().astype(np.uint64)
        assert int(d_status[0]) == 0
        if nsl % 8 == 0:                                # oracle check of one chunk of every 8th slice
            c = (nsl // 8) % (len(off) - 1)
            want = O.encode_chunk(x[int(off[c]):int(off[c + 1])].cpu().numpy(), M, L)
            got = comp[pos + int(b[c]):pos + int(b[c + 1])].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want), "c4: stream differs from the oracle"
            checked += 1
        boffs.append(b[1:] + np.uint64(pos))
        pos += int(b[-1])
        nsl += 1
        last_x = x
    boff = np.concatenate(boffs)
    nchunks = len(boff) - 1
    n_waves = nchunks * wpc
    off_all = d.chunk_offsets(wpc * L, n_waves * L)
    y = torch.empty(n_waves * L, dtype=torch.int16, device=dev)
    stream = comp[:pos]
    codec.decode_device_async(stream, boff, off_all, M, L, y, d_status)
    torch.cuda.synchronize()
    assert int(d_status[0]) == 0 and torch.equal(y[-last_x.numel():], last_x), "c4: decode(encode(x)) != x"
    return stream, boff, off_all, y, M, L, n_waves, checked


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
    my_cores = pin_rank_to_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import deltarice_b200 as d
    from deltarice_b200 import shard
    from oracle import oracle as O

    name = args.workload
    scaling = WORKLOADS[name][5]
    codec = d.DeltaRice(local_rank)
    d_status = torch.zeros(2, dtype=torch.int32, device=dev)
    props = torch.cuda.get_device_properties(dev)
    uuid = "GPU-" + str(props.uuid) if not str(props.uuid).startswith("GPU-") else str(props.uuid)
    clocks = ClockSampler(uuid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    first_chunk = 0
    decode_only = name == "c4"
    if decode_only:
        comp, boff, off, y, M, L, n_waves_all, checked = run_c4(args, dev, codec)
        x = out = d_boff = None
        raw_bytes = y.numel() * 2
        comp_bytes = comp.numel()
        nchunks = len(boff) - 1
        n_waves_local = n_waves_all
        extra["c4"] = {"compressed_gib": round(comp_bytes / GIB, 3), "raw_gib": round(raw_bytes / GIB, 3),
                       "chunks": nchunks, "oracle_checked_chunks": checked,
                       "note": "one drice_decode_batch_dev_async call over the whole stream"}
    else:
        x, off, M, L, n_waves_all, first_chunk = make_input(name, rank, world, dev, args)
        raw_bytes = x.numel() * 2
        nchunks = len(off) - 1
        n_waves_local = x.numel() // L
        big = raw_bytes > 16 * GIB                      # c5: no room for the worst-case bound (capacity errors are reported)
        cap = int(raw_bytes * 0.55) + (1 << 20) if big else codec.bound_bytes(off, L)
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_boff = torch.zeros(nchunks + 1, dtype=torch.int64, device=dev)
        y = torch.empty_like(x)
    ms_list = [M] if name != "c3" else list(C3_SWEEP)
    sweep = []
    gathered = {}

    for Mi in ms_list:
        if not decode_only:
            # one checked pass: offsets for the decode calls, round trip, status
            codec.encode_device_async(x, off, Mi, L, out, d_boff, d_status)
            boff = d_boff.cpu().numpy().astype(np.uint64)
            assert int(d_status[0]) == 0, "encode status"
            comp_bytes = int(boff[-1])
            comp = out[:comp_bytes]
            y.zero_()
            codec.decode_device_async(comp, boff, off, Mi, L, y, d_status)
            torch.cuda.synchronize()
            assert int(d_status[0]) == 0 and torch.equal(x, y), "decode(encode(x)) != x"
            if name in ("c3", "c5"):                    # bit-exact against the oracle on sampled chunks
                for c in sorted({0, nchunks // 2, nchunks - 1}):
                    want = O.encode_chunk(x[int(off[c]):int(off[c + 1])].cpu().numpy(), Mi, L)
                    got = comp[int(boff[c]):int(boff[c + 1])].cpu().numpy().view(np.uint32)
                    assert np.array_equal(got, want), f"{name}: chunk {c} differs from the oracle (M={Mi})"
        ratio = comp_bytes / raw_bytes

        # the path's one exchange (per-shard byte counts -> offsets in the concatenated stream) needs the
        # encoder's result but nothing of the decoder: it runs on a side stream next to the decode kernels
        # and is joined before the step ends
        side = torch.cuda.Stream(device=dev) if world > 1 else None

        def exchange():
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gathered["counts"], gathered["offsets"] = shard.gather_shard_offsets(d_boff[nchunks:nchunks + 1])

        def step(ev=None, s=0):
            if not decode_only:
                codec.encode_device_async(x, off, Mi, L, out, d_boff, d_status)
            if ev is not None:
                ev[2 * s + 1].record()
            if world > 1 and not decode_only:
                exchange()
            codec.decode_device_async(comp, boff, off, Mi, L, y, d_status)
            if world > 1 and not decode_only:
                torch.cuda.current_stream(dev).wait_stream(side)
            if ev is not None:
                ev[2 * s + 2].record()

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
            step(ev, s)
        barrier()
        clocks.live(False)
        launches = codec.launches - l0
        t_total = ev[0].elapsed_time(ev[-1])                       # ms, device clock
        t_enc = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps))
        t_dec = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps))
        ktimes = codec.timing_read(reset=True)
        codec.timing(False)
        assert int(d_status[0]) == 0
        tt = torch.tensor([t_total, t_enc, t_dec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total, t_enc, t_dec = [float(v) for v in tt.cpu()]
        kern = {k: {"launches": c, "avg_ms": round(ms / c, 4)} for k, (ms, c) in ktimes.items() if c}
        sweep.append({"M": Mi, "ratio": round(ratio, 5), "t_total": t_total, "t_enc": t_enc, "t_dec": t_dec,
                      "kern": kern, "launches": launches, "comp_bytes": comp_bytes})

    # ---- c5: the shards at their gathered offsets = ONE stream; check it (outside the timed region) ----
    parity = None
    if name == "c5":
        counts = gathered["counts"].cpu().numpy() if world > 1 else np.array([comp_bytes])
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        assert int(counts[rank]) == comp_bytes
        wpc = WORKLOADS[name][3]
        nchunks_all = n_waves_all // wpc
        # every rank's chunk offsets, rebased into the one stream
        g_boff = torch.zeros(nchunks_all + 1, dtype=torch.int64, device=dev)
        mine = torch.from_numpy(shard.global_chunk_byte_offsets(boff[:-1], offsets, rank).astype(np.int64)).to(dev)
        g_boff[first_chunk:first_chunk + nchunks] = mine
        if world > 1:
            dist.all_reduce(g_boff, op=dist.ReduceOp.SUM)
        g_boff[nchunks_all] = int(offsets[-1])
        stream = None
        if rank == 0:
            stream = torch.empty(int(offsets[-1]), dtype=torch.uint8, device=dev)
            stream[:comp_bytes] = comp
            for r in range(1, world):
                dist.recv(stream[int(offsets[r]):int(offsets[r + 1])], src=r)
        else:
            dist.send(comp.contiguous(), dst=0)
        if rank == 0:
            gb = g_boff.cpu().numpy()
            # (1) host-side walk: every chunk starts with its sample count where the offsets say it does
            heads = stream.view(torch.int32)[torch.from_numpy(gb[:-1] // 4).to(dev)].cpu().numpy()
            walk_ok = bool(np.all(heads == wpc * L)) and bool(np.all(np.diff(gb) > 0))
            # (2) sampled chunks, regenerated from their seeds, against the oracle's stream
            from deltarice_b200 import synth
            sample = sorted({0, 1, nchunks_all // world, nchunks_all // 2, nchunks_all - 1} |
                            {shard.shard_chunk_range(nchunks_all, world, r)[0] for r in range(world)})
            sample = [c for c in sample if c < nchunks_all]
            ok = True
            for c in sample:
                raw_c = synth.nab_like_torch(wpc, L, SEED + c, dev).reshape(-1).cpu().numpy()
                want = O.encode_chunk(raw_c, M, L)
                got = stream[int(gb[c]):int(gb[c + 1])].cpu().numpy().view(np.uint32)
                ok = ok and np.array_equal(got, want)
            parity = {"concatenated_stream_bytes": int(offsets[-1]), "header_walk_ok": walk_ok,
                      "chunks_compared_with_oracle": len(sample), "bit_exact": bool(ok),
                      "how": "shards sent to rank 0 (NCCL send/recv) and placed at the all-gathered offsets; "
                             "host walk of all chunk headers + sampled chunks against the oracle"}
            assert walk_ok and ok, "c5: concatenated stream differs from the single-stream oracle"
            del stream
        barrier()

    # ---- e2e: host buffers through the chunk scheduler (what the H5Z callback calls) ----
    # (a) sequential: encode the batch, then decode its streams (one handle, one direction of
    #     PCIe busy at a time);
    # (b) pipelined (the headline): two handles on two host threads - step s encodes the batch into
    #     stream buffer s%2 while the streams step s-1 produced are decoded, so raw-in and raw-out
    #     share the link's two directions.  Every step still encodes one batch and decodes one.
    # (c) h5z: H5Z_filter_deltarice per chunk with malloc'ed buffers (the reference arm's call pattern).
    e2e = None
    if (name in ("c2", "c1") or args.force_e2e) and not decode_only:
        M0 = ms_list[0]
        e2e_steps = max(1, args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps))
        cs = WORKLOADS[name][3] * L
        nx = min(x.numel(), 153391 * 3500)              # at most ~1 GB of this rank's samples
        if nx >= cs:
            nx -= nx % cs if x.numel() > nx else 0
        xs = x[:nx]
        offs = d.chunk_offsets(min(cs, nx), nx)
        ncs = len(offs) - 1
        caps = codec.bound_bytes(offs, L)
        h_raw = codec.pinned_empty(nx, np.int16)
        h_raw[:] = xs.cpu().numpy()
        h_comp = [codec.pinned_empty(caps, np.uint8) for _ in range(2)]
        h_back = codec.pinned_empty(nx, np.int16)
        h_boff = [np.zeros(ncs + 1, dtype=np.uint64) for _ in range(2)]
        nb = 0
        for i in range(2):
            nb = codec.encode_host_into(h_raw, offs, M0, L, h_comp[i], h_boff[i])
            codec.decode_host_into(h_comp[i][:nb], h_boff[i], offs, M0, L, h_back)
        assert np.array_equal(h_back, h_raw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nb = codec.encode_host_into(h_raw, offs, M0, L, h_comp[0], h_boff[0])
            codec.decode_host_into(h_comp[0][:nb], h_boff[0], offs, M0, L, h_back)
        torch.cuda.synchronize()
        t_e2e_seq = (time.perf_counter() - t0) * 1e3 / e2e_steps   # ms per step (wall: host work is part of it)

        codec2 = d.DeltaRice(local_rank)
        h_back[:] = 0

        def _enc(i):
            codec.encode_host_into(h_raw, offs, M0, L, h_comp[i], h_boff[i])

        def _dec(i):
            codec2.decode_host_into(h_comp[i][:nb], h_boff[i], offs, M0, L, h_back)

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

        # (c) the H5Z callback, chunk by chunk, malloc'ed buffers: same chunks and call pattern as --impl reference
        from deltarice_b200 import _lib
        fr = FilterRunner()
        lib = _lib.load()
        os.environ.setdefault("DRICE_DEVICE", str(local_rank))
        hchunks = [np.ascontiguousarray(h_raw[int(offs[c]):int(offs[c + 1])]) for c in range(min(ncs, args.ref_chunks))]
        hraw = sum(c.nbytes for c in hchunks)
        fr.step(lib, hchunks, L, M0)
        barrier()
        te = td = 0.0
        for _ in range(e2e_steps):
            a, b, _ = fr.step(lib, hchunks, L, M0)
            te += a
            td += b
        t_h5z = (te + td) * 1e3 / e2e_steps
        e2e_raw = nx * 2
        tt = torch.tensor([t_e2e, t_e2e_seq, t_h5z, te * 1e3 / e2e_steps, td * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e, t_e2e_seq, t_h5z, t_h5z_e, t_h5z_d = [float(v) for v in tt.cpu()]
        h2d_gbs, d2h_gbs = pcie_duplex_gbs(dev, barrier)
        hbt = torch.tensor([host_copy_gbs(), h2d_gbs, d2h_gbs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(hbt, op=dist.ReduceOp.SUM)
        # a step moves (raw + stream) bytes each way: the raw-int16 rate the measured link rates allow
        link_ceiling = 2 * e2e_raw / ((e2e_raw + nb) / min(float(hbt[1]), float(hbt[2])))
        e2e = {
            "value": round(world * 2 * e2e_raw / (t_e2e * 1e6), 2), "unit": UNIT,
            "h2d_bytes_per_step": e2e_raw + nb, "d2h_bytes_per_step": nb + e2e_raw,
            "steps": e2e_steps, "ms_per_step": round(t_e2e, 2),
            "api": "drice_encode_batch_host + drice_decode_batch_host (pinned host buffers), two handles: "
                   "step s encodes while the streams of step s-1 are decoded (both PCIe directions busy)",
            "sequential": {"value": round(world * 2 * e2e_raw / (t_e2e_seq * 1e6), 2), "ms_per_step": round(t_e2e_seq, 2),
                           "api": "one handle: encode the batch, then decode its streams"},
            "h5z": {"value": round(world * 2 * hraw / (t_h5z * 1e6), 3), "unit": UNIT, "ms_per_step": round(t_h5z, 2),
                    "encode_gbs": round(world * hraw / (t_h5z_e * 1e6), 3), "decode_gbs": round(world * hraw / (t_h5z_d * 1e6), 3),
                    "chunks_per_step": len(hchunks), "chunk_bytes": int(hchunks[0].nbytes), "steps": e2e_steps,
                    "api": "H5Z_filter_deltarice(flags, 2, {M, L}, nbytes, &buf_size, &buf): one call per chunk, "
                           "malloc'ed pageable buffers, ownership handed over as libhdf5 does; timed around the calls "
                           "(same chunks, call pattern and timing as --impl reference)"},
            "host": {"cores_per_rank": len(my_cores), "memcpy_gbs_all_ranks": round(float(hbt[0]), 1),
                     "h2d_gbs_all_ranks": round(float(hbt[1]), 1), "d2h_gbs_all_ranks": round(float(hbt[2]), 1),
                     "e2e_ceiling_gbs": round(link_ceiling, 1), "e2e_frac_of_ceiling": round(world * 2 * e2e_raw / (t_e2e * 1e6) / link_ceiling, 3),
                     "note": "measured in this run, every rank at once: numpy memcpy per rank (summed), pinned host<->device "
                             "copies with both directions busy (summed); e2e_ceiling = raw int16 GB/s (both directions counted, "
                             "as `value`) if the links ran at those rates and nothing else took time"},
        }
    clk = clocks.result()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        raw_all = raw_bytes * world if scaling == "weak" else n_waves_all * L * 2
        for sw in sweep:
            ms_step = sw["t_total"] / args.steps
            alg = raw_bytes * (1.0 + sw["ratio"])          # per rank: read raw + write stream (encode) / the reverse
            sw["value"] = round((1 if decode_only else 2) * raw_all / (ms_step * 1e6), 1)
            sw["ms_per_step"] = round(ms_step, 4)
            sw["encode_gbs"] = None if decode_only else round(raw_all * args.steps / (sw["t_enc"] * 1e6), 1)
            sw["decode_gbs"] = round(raw_all * args.steps / (sw["t_dec"] * 1e6), 1)
            for k, v in sw["kern"].items():
                a = 4.0 * (nchunks + n_waves_local) if k == "locate_kernel" else alg
                v["achieved_gbs"] = round(a / (v["avg_ms"] * 1e6), 1)
                v["frac_of_peak"] = round(a / (v["avg_ms"] * 1e6) / peak, 4)
            dec_ms = sum(v["avg_ms"] for k, v in sw["kern"].items() if k != "encode_kernel")
            if dec_ms:
                sw["decode_frac_of_peak"] = round(alg / (dec_ms * 1e6) / peak, 4)
        main = sweep[len(sweep) // 2] if name == "c3" else sweep[0]
        kern = main["kern"]
        dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches"]) if kern else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom) if name == "c2" else None
        except Exception:  # noqa: BLE001
            pass
        roof = None
        if dom:
            alg = raw_bytes * (1.0 + main["ratio"])
            alg_dom = 4.0 * (nchunks + n_waves_local) if dom == "locate_kernel" else alg
            ach = alg_dom / (kern[dom]["avg_ms"] * 1e6)
            roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(alg_dom), "all_kernels": kern}
            if "decode_frac_of_peak" in main:
                roof["decode_frac"] = main["decode_frac_of_peak"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline and name in ("c2", "c3", "c5", "c1"):
            try:
                allc = cpu_baseline_subprocess(name, False, args.ref_chunks, 3, 1)
                one = cpu_baseline_subprocess(name, True, max(1, args.ref_chunks // 4), 1, 1)
                cpu = dict(allc["cpu_baseline"])
                cpu["encode_gbs"], cpu["decode_gbs"] = allc["encode_gbs"], allc["decode_gbs"]
                cpu["single_thread"] = {"value": one["value"], "encode_gbs": one["encode_gbs"],
                                        "decode_gbs": one["decode_gbs"], "cores": 1, "kind": one["cpu_baseline"]["kind"]}
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: " + repr(e)[:200]}
        cfg = workload_config(name, world, n_waves_all if (name == "c4" or args.c5_chunks) else None)
        line = {
            "metric": METRIC if not decode_only else "decode GB/s of raw int16 (bit-exact Delta-Rice stream, decode only)",
            "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "int16/u32", "data": "synthetic", "config": cfg,
            "encode_gbs": main["encode_gbs"], "decode_gbs": main["decode_gbs"], "ratio": main["ratio"],
            "e2e": e2e, "gpu_launches": int(main["launches"]), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        if name == "c3":
            line["sweep"] = [{k: sw[k] for k in ("M", "ratio", "value", "ms_per_step", "encode_gbs", "decode_gbs", "kern",
                                                  "decode_frac_of_peak") if k in sw} for sw in sweep]
        if parity is not None:
            line["parity"] = parity
        line.update(extra)
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
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e legs (0 = --steps)")
    ap.add_argument("--force-e2e", action="store_true", help="run the e2e legs for workloads other than c2")
    ap.add_argument("--ref-chunks", type=int, default=8, help="chunks per step of the CPU reference sample / the e2e.h5z leg")
    ap.add_argument("--ref-serial", action="store_true", help="reference arm single-threaded (serial build)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-gib", type=float, default=32.0, help="c4: GiB of compressed stream to decode")
    ap.add_argument("--c5-chunks", type=int, default=0, help="c5: chunks of the whole dataset (0 = 2454 = 64 GiB)")
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
