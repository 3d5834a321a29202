"""Per-chunk latency of the H5Z filter callback itself (one chunk per call, malloc'ed buffers handed
over exactly as libhdf5 does; only the callback is timed) against the unmodified reference on the host
cores: README config C1 chunks (20 x 7000 int16 = 280 KB) and a 28 MB chunk (2000 x 7000)."""
import os, sys, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltarice_b200 import _lib
from oracle import oracle as O
libc = C.CDLL(None); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]

def call(lib, data, cd, reverse):
    n = len(data); p = libc.malloc(n + 64); C.memmove(p, data, n)
    buf, bs = C.c_void_p(p), C.c_size_t(n); cdv = (C.c_uint * len(cd))(*cd)
    t0 = time.perf_counter()
    ret = lib.H5Z_filter_deltarice(0x100 if reverse else 0, len(cd), cdv, n, C.byref(bs), C.byref(buf))
    dt = time.perf_counter() - t0
    assert ret not in (0, C.c_size_t(-1).value)
    out = C.string_at(buf.value, ret); libc.free(buf)
    return out, dt

def bench(lib, raw, cd, n):
    s, _ = call(lib, raw, cd, False); call(lib, s, cd, True)
    te = sum(call(lib, raw, cd, False)[1] for _ in range(n)) / n
    td = sum(call(lib, s, cd, True)[1] for _ in range(n)) / n
    return 1e3 * te, 1e3 * td, s

ours = _lib.load()
ref = O.ref_lib("omp") if O.ref_available("omp") else None
rng = np.random.default_rng(0)
for rows in (20, 200, 2000):
    raw = rng.normal(0, 10, (rows, 7000)).astype(np.int16).tobytes(); cd = (8, 7000); n = 30 if rows <= 200 else 5
    te, td, s = bench(ours, raw, cd, n)
    line = f"chunk {rows}x7000 ({len(raw)/1e6:.2f} MB): GPU filter encode {te:.3f} ms, decode {td:.3f} ms"
    if ref is not None:
        re_, rd, rs = bench(ref, raw, cd, n); assert rs == s
        line += f" | reference (OpenMP, {len(os.sched_getaffinity(0))} cores) encode {re_:.3f} ms, decode {rd:.3f} ms"
    print(line)
