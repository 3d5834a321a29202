"""decode time vs number of waves, wide (CTA per wave) vs lane (lane per wave) parser: run twice with
DRICE_PARSE_WIDE=1000000 and DRICE_PARSE_WIDE=0"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deltarice_b200 as d
from deltarice_b200.synth import nab_like_torch
codec = d.DeltaRice(0)
for L, M in ((3500, 4), (7000, 8)):
    for nw in (500, 1000, 2000, 4000, 8000, 16000, 32000):
        x = nab_like_torch(nw, L, 1, "cuda").reshape(-1)
        off = d.chunk_offsets(min(2000, nw) * L, x.numel())
        comp, boff = codec.encode_device(x, off, M, L)
        y = torch.empty_like(x); st = torch.zeros(2, dtype=torch.int32, device="cuda")
        for _ in range(3): codec.decode_device_async(comp, boff, off, M, L, y, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); codec.decode_device_async(comp, boff, off, M, L, y, st); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert torch.equal(x, y)
        print(f"L={L} waves={nw:6d} decode {np.median(ts):.3f} ms")
