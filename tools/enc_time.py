"""Encode-only kernel timing (no correctness check): python tools/enc_time.py [n_waves L M wpc reps]"""
import os, sys
import numpy as np, torch
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
out = torch.empty(codec.bound_bytes(off, L), dtype=torch.uint8, device="cuda")
d_boff = torch.zeros(len(off), dtype=torch.int64, device="cuda")
d_status = torch.zeros(2, dtype=torch.int32, device="cuda")
for _ in range(4):
    codec.encode_device_async(x, off, M, L, out, d_boff, d_status)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); codec.encode_device_async(x, off, M, L, out, d_boff, d_status); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"dbg={os.environ.get('DRICE_ENC_SEG_DBG','0')} encode median {np.median(ts):.3f} ms best {min(ts):.3f} (incl. prep kernel) bytes {int(d_boff[-1])}")
