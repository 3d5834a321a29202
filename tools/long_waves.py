"""Long waves (reference docs/Performance.md:27-47: nEDM 32 x 81920, NOPTREX 32 x 500000, and the default
WaveformLength = -1: the whole chunk is ONE wave): kernel times on device-resident data, the H5Z callback
per chunk, and the unmodified reference on the host cores beside it.
usage: python tools/long_waves.py [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deltarice_b200 as d
from deltarice_b200 import h5, synth
from oracle import oracle as O

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
codec = d.DeltaRice(0)
os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))


def ev_time(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def wall(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


for name, nw, L, M, Lopt in (("nEDM-like 32 x 81920", 32, 81920, 8, 81920), ("NOPTREX-like 32 x 500000", 32, 500000, 8, 500000),
                             ("whole chunk one wave (L=-1) 2000 x 7000", 1, 14000000, 8, None)):
    x = synth.nab_like(nw, L, seed=11).ravel() if L <= 500000 else synth.nab_like(2000, 7000, seed=11).ravel()
    raw = x.nbytes
    off = np.array([0, x.size], dtype=np.uint64)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty(codec.bound_bytes(off, Lopt), dtype=torch.uint8, device="cuda")
    d_boff = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_status = torch.zeros(2, dtype=torch.int32, device="cuda")
    te = ev_time(lambda: codec.encode_device_async(xd, off, M, Lopt, out, d_boff, d_status), reps)
    boff = d_boff.cpu().numpy().astype(np.uint64)
    nb = int(boff[-1])
    want = O.encode_chunk(x, M, Lopt)
    exact = bool(np.array_equal(out[:nb].cpu().numpy().view(np.uint32), want))
    y = torch.empty_like(xd)
    td = ev_time(lambda: codec.decode_device_async(out[:nb], boff, off, M, Lopt, y, d_status), reps)
    rt = bool(torch.equal(xd, y)) and int(d_status[0]) == 0
    cd = (M,) if Lopt is None else (M, Lopt)
    from bench import FilterRunner
    from deltarice_b200 import _lib
    fr = FilterRunner()
    ours = _lib.load()
    s, _ = fr.call(ours, x, cd, False)                      # (timed around the callback only: malloc'ed buffers as libhdf5 hands them)
    fr.call(ours, s, cd, True)
    he = float(np.median([fr.call(ours, x, cd, False)[1] for _ in range(reps)])) * 1e3
    hd = float(np.median([fr.call(ours, s, cd, True)[1] for _ in range(reps)])) * 1e3
    line = (f"{name}: raw {raw / 1e6:.1f} MB ratio {nb / raw:.3f} bit-exact {exact} roundtrip {rt} | kernels encode {te:.3f} ms "
            f"({raw / te / 1e6:.1f} GB/s) decode {td:.3f} ms ({raw / td / 1e6:.1f} GB/s) | H5Z callback encode {he:.2f} ms decode {hd:.2f} ms")
    if O.ref_available("omp"):
        lib = O.ref_lib("omp")
        cdv = (M, Lopt) if Lopt is not None else (M,)
        def ref_enc():
            return fr.call(lib, x, cdv, False)
        sref, _ = ref_enc()
        re_ = float(np.median([ref_enc()[1] for _ in range(reps)])) * 1e3
        rd_ = float(np.median([fr.call(lib, sref, cdv, True)[1] for _ in range(reps)])) * 1e3
        line += f" | reference ({os.environ['OMP_NUM_THREADS']} threads) encode {re_:.2f} ms decode {rd_:.2f} ms"
    print(line, flush=True)
