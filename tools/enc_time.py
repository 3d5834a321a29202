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
nb = int(d_boff[-1])
ok = int(d_status[0]) == 0
# checksum of the stream (compare across variants) and the round trip
w = out[:nb].view(torch.int32).to(torch.int64)
chk = int((w * (torch.arange(w.numel(), device="cuda") % 65521 + 1)).sum())
y = torch.empty_like(x)
codec.decode_device_async(out[:nb], d_boff.cpu().numpy().astype(np.uint64), off, M, L, y, d_status)
torch.cuda.synchronize()
rt = bool(torch.equal(x, y)) and int(d_status[0]) == 0
print(f"lut={os.environ.get('DRICE_ENC_LUT','-')} waves={n_waves} L={L} M={M} encode median {np.median(ts):.3f} ms best {min(ts):.3f} (incl. prep kernel) "
      f"bytes {nb} chk {chk} status_ok {ok} roundtrip {rt}")
