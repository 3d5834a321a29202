"""PCIe duplex check: host-pointer encode / decode alone and concurrently, plus raw copies."""
import sys, time, threading, numpy as np, torch
sys.path.insert(0, '/root/repo')
import deltarice_b200 as d
from deltarice_b200.synth import nab_like_torch
nw, L, M, wpc = 153391, 3500, 4, 2000
x = nab_like_torch(nw, L, 20251018, "cuda").reshape(-1)
off = d.chunk_offsets(wpc * L, x.numel())
n = len(off) - 1
c1, c2 = d.DeltaRice(0), d.DeltaRice(0)
cap = c1.bound_bytes(off, L)
h_raw = c1.pinned_empty(x.numel(), np.int16); h_raw[:] = x.cpu().numpy()
h_comp = c1.pinned_empty(cap, np.uint8); h_back = c1.pinned_empty(x.numel(), np.int16)
h_boff = np.zeros(n + 1, dtype=np.uint64)
nb = c1.encode_host_into(h_raw, off, M, L, h_comp, h_boff)
def enc(): c1.encode_host_into(h_raw, off, M, L, h_comp, h_boff.copy())
def dec(): c2.decode_host_into(h_comp[:nb], h_boff, off, M, L, h_back)
def timeit(fs, reps=4):
    for f in fs: f()
    t0 = time.perf_counter()
    for _ in range(reps):
        th = [threading.Thread(target=f) for f in fs]
        [t.start() for t in th]; [t.join() for t in th]
    return (time.perf_counter() - t0) / reps * 1e3
print("enc alone ms", timeit([enc])); print("dec alone ms", timeit([dec])); print("both ms", timeit([enc, dec]))
# raw copies
a = torch.empty(2**30, dtype=torch.uint8, device="cuda"); b = torch.empty(2**30, dtype=torch.uint8, device="cuda")
ha = torch.empty(2**30, dtype=torch.uint8).pin_memory(); hb = torch.empty(2**30, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def cp(h2d, d2h, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): a.copy_(ha, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hb.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
cp(1, 1)
print("1 GiB h2d ms", cp(1, 0), "d2h ms", cp(0, 1), "both ms", cp(1, 1))
