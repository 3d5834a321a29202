"""Randomised parity runs against the oracle (bit-exact stream, exact round trip, oracle stream decodes):
python tools/fuzz_gpu.py [iterations] [seed].  Shapes cover tiny and mid-size batches, ragged chunks, every
Rice parameter class, long waves, whole-chunk waves, mixed noise levels (escape-heavy waves next to quiet ones).
Run it under `timeout` on the GPU box: a kernel that hangs would otherwise hold the box until its limit."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deltarice_b200 as d
from oracle import oracle as O

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
r = np.random.default_rng(seed)
codec = d.DeltaRice(0)
t0 = time.time()
for it in range(iters):
    kind = r.integers(0, 6)
    if kind == 0:   L = int(r.integers(1, 40))
    elif kind == 1: L = int(r.integers(40, 2000))
    elif kind == 2: L = int(r.integers(2000, 8193))
    elif kind == 3: L = int(r.integers(8193, 60000))
    elif kind == 4: L = None
    else:           L = int(r.choice([3500, 7000, 1024, 8192, 512]))
    M = int(2 ** r.choice([0, 1, 2, 3, 3, 3, 4, 5, 6, 8, 11]))
    nch = int(r.integers(1, 7))
    Lw = L if L else int(r.integers(1, 30000))
    budget = 3_000_000
    wpc_max = max(1, min(3000 if r.random() < 0.15 else 300, budget // (Lw * nch)))
    sizes = []
    for c in range(nch):
        w = int(r.integers(0, wpc_max + 1)) if r.random() < 0.9 else 0
        n = w * Lw + (int(r.integers(0, Lw)) if r.random() < 0.4 else 0)
        if L is None: n = int(r.integers(0, 30000))
        sizes.append(n)
    if sum(sizes) == 0: sizes[0] = 1
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    tot = int(off[-1])
    nwv = max(1, tot // max(Lw, 1) + nch)
    sig = r.choice([0.5, 2, 5, 10, 40, 300, 5000], size=nwv)
    if r.random() < 0.5: sig[:] = sig[0]
    s_per = np.repeat(sig, Lw)[:tot] if tot <= nwv * Lw else np.resize(np.repeat(sig, Lw), tot)
    if M == 1: s_per = np.minimum(s_per, 1.5)           # (the reference hangs on M = 1 with large deltas; ours escapes - keep both happy)
    base = np.cumsum(r.normal(0, 1, tot) * s_per * 0.2) if r.random() < 0.5 else 0
    x = np.clip(np.rint(r.normal(0, 1, tot) * s_per + base), -32768, 32767).astype(np.int16)
    parts, boff = [], [0]
    for c in range(nch):
        s = O.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L)
        parts.append(s); boff.append(boff[-1] + 4 * s.size)
    want = np.concatenate(parts); wboff = np.array(boff, dtype=np.uint64)
    got, gb = codec.encode_host(x, off, M, L)
    ok = np.array_equal(gb, wboff) and np.array_equal(got.view(np.uint32), want)
    ok = ok and np.array_equal(codec.decode_host(want.view(np.uint8), wboff, off, M, L), x)
    if not ok:
        print(f"MISMATCH it={it} seed={seed} L={L} M={M} sizes={sizes}")
        sys.exit(1)
print(f"fuzz ok: {iters} cases, seed {seed}, {time.time() - t0:.1f} s, env " +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DRICE_")))
