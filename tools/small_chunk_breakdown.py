"""Where the time of ONE small chunk (20 x 7000, the README configuration) goes: kernels only (device-resident,
enqueue + synchronize), the host-pointer batch call with pinned buffers, the H5Z callback with malloc'ed buffers."""
import os, sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deltarice_b200 as d
from deltarice_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L, M, n = 7000, 8, 200
x = np.random.default_rng(0).normal(0, 10, (rows, L)).astype(np.int16).ravel()
codec = d.DeltaRice(0)
off = np.array([0, x.size], dtype=np.uint64)
xd = torch.from_numpy(x).cuda()
out = torch.empty(codec.bound_bytes(off, L), dtype=torch.uint8, device="cuda")
boff = torch.zeros(2, dtype=torch.int64, device="cuda"); st = torch.zeros(2, dtype=torch.int32, device="cuda")
def t(f, n=n):
    for _ in range(10): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
def dev_enc():
    codec.encode_device_async(xd, off, M, L, out, boff, st); torch.cuda.synchronize()
print(f"{rows} x {L}: device-resident encode, enqueue + synchronize: {t(dev_enc):.1f} us")
nb = int(boff[1]); comp_d = out[:nb].clone(); y = torch.empty_like(xd); bo = np.array([0, nb], dtype=np.uint64)
def dev_dec():
    codec.decode_device_async(comp_d, bo, off, M, L, y, st); torch.cuda.synchronize()
print(f"device-resident decode: {t(dev_dec):.1f} us")
px = codec.pinned_empty(x.size, np.int16); px[:] = x
pout = codec.pinned_empty(codec.bound_bytes(off, L), np.uint8); pb = np.zeros(2, dtype=np.uint64)
print(f"encode_host_into, pinned in / out: {t(lambda: codec.encode_host_into(px, off, M, L, pout, pb)):.1f} us")
pc = codec.pinned_empty(nb, np.uint8); pc[:] = pout[:nb]; py = codec.pinned_empty(x.size, np.int16)
print(f"decode_host_into, pinned in / out: {t(lambda: codec.decode_host_into(pc, bo, off, M, L, py)):.1f} us")
xs = x.copy(); outp = np.empty(codec.bound_bytes(off, L), np.uint8)
print(f"encode_host_into, pageable in / out: {t(lambda: codec.encode_host_into(xs, off, M, L, outp, pb)):.1f} us")
lib = _lib.load(); libc = C.CDLL(None); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
raw = x.tobytes(); cdv = (C.c_uint * 2)(M, L)
def h5z(data, rev):
    nby = len(data); p = libc.malloc(nby + 64); C.memmove(p, data, nby)
    buf, bs = C.c_void_p(p), C.c_size_t(nby)
    t0 = time.perf_counter()
    ret = lib.H5Z_filter_deltarice(0x100 if rev else 0, 2, cdv, nby, C.byref(bs), C.byref(buf))
    dt = time.perf_counter() - t0
    s = C.string_at(buf.value, ret); libc.free(buf)
    return s, dt
s, _ = h5z(raw, False)
for _ in range(10): h5z(raw, False); h5z(s, True)
print(f"H5Z_filter_deltarice encode: {np.mean([h5z(raw, False)[1] for _ in range(n)]) * 1e6:.1f} us, decode: {np.mean([h5z(s, True)[1] for _ in range(n)]) * 1e6:.1f} us")
codec.timing(True)
for _ in range(20): dev_enc(); dev_dec()
print("kernel times (CUDA events around each launch group, ms):", {k: (round(v["avg_ms"] * 1e3, 1) if isinstance(v, dict) and "avg_ms" in v else v) for k, v in codec.timing_read().items()})
