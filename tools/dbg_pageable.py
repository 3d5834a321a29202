"""Host-pointer entry points with pageable vs pinned buffers (one 28 MB chunk, as H5Z hands it over)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deltarice_b200 as d
rng = np.random.default_rng(0)
x = rng.normal(0, 10, (2000, 7000)).astype(np.int16).ravel()
off = np.array([0, x.size], dtype=np.uint64)
c = d.DeltaRice(0)
cap = c.bound_bytes(off, 7000)
def t(fn, n=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
# pageable in / pageable out
out_pg = np.empty(cap, np.uint8); boff = np.zeros(2, np.uint64)
print("encode pageable->pageable %.2f ms" % t(lambda: c.encode_host_into(x, off, 8, 7000, out_pg, boff)))
nb = int(boff[1])
xp = c.pinned_empty(x.size, np.int16); xp[:] = x
out_pin = c.pinned_empty(cap, np.uint8)
print("encode pinned->pinned     %.2f ms" % t(lambda: c.encode_host_into(xp, off, 8, 7000, out_pin, boff)))
print("encode pageable->pinned   %.2f ms" % t(lambda: c.encode_host_into(x, off, 8, 7000, out_pin, boff)))
print("encode pinned->pageable   %.2f ms" % t(lambda: c.encode_host_into(xp, off, 8, 7000, out_pg, boff)))
comp_pg = out_pg[:nb].copy(); back_pg = np.empty_like(x)
comp_pin = c.pinned_empty(nb, np.uint8); comp_pin[:] = comp_pg; back_pin = c.pinned_empty(x.size, np.int16)
print("decode pageable->pageable %.2f ms" % t(lambda: c.decode_host_into(comp_pg, boff, off, 8, 7000, back_pg)))
print("decode pinned->pinned     %.2f ms" % t(lambda: c.decode_host_into(comp_pin, boff, off, 8, 7000, back_pin)))
print("memcpy 28 MB host->host   %.2f ms" % t(lambda: np.copyto(back_pg, x)))
