"""Quick kernel timing (CUDA events, data resident in HBM): encode / decode GB/s of raw int16.
usage: python tools/quick_bench.py [n_waves] [L] [M] [waves_per_chunk] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deltarice_b200 as d
from deltarice_b200.synth import nab_like_torch

n_waves = int(sys.argv[1]) if len(sys.argv) > 1 else 153391
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3500
M = int(sys.argv[3]) if len(sys.argv) > 3 else 4
wpc = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10

codec = d.DeltaRice(0)
x = nab_like_torch(n_waves, L, 20251018, "cuda").reshape(-1)
off = d.chunk_offsets(wpc * L, x.numel())
n = len(off) - 1
cap = codec.bound_bytes(off, L)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_boff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
d_status = torch.zeros(2, dtype=torch.int32, device="cuda")
y = torch.empty_like(x)
raw_bytes = x.numel() * 2


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.array(ts)


te = timeit(lambda: codec.encode_device_async(x, off, M, L, out, d_boff, d_status), reps)
boff = d_boff.cpu().numpy().astype(np.uint64)
comp_bytes = int(boff[-1])
assert int(d_status[0]) == 0
comp = out[:comp_bytes]
td = timeit(lambda: codec.decode_device_async(comp, boff, off, M, L, y, d_status), reps)
assert int(d_status[0]) == 0
assert torch.equal(x, y)
ratio = comp_bytes / raw_bytes
alg = raw_bytes * (1 + ratio)
print(f"waves={n_waves} L={L} M={M} chunks={n} raw={raw_bytes/1e9:.3f} GB ratio={ratio:.4f}")
for nm, t in (("encode", te), ("decode", td)):
    print(f"{nm}: median {np.median(t):.3f} ms  best {t.min():.3f} ms  raw {raw_bytes/np.median(t)/1e6:.1f} GB/s  "
          f"alg {alg/np.median(t)/1e6:.1f} GB/s  frac_of_6456.8 {alg/np.median(t)/1e6/6456.8:.3f}")
