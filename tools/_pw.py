import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import deltarice_b200 as d
rows=20; L=7000; M=8
x = np.random.default_rng(0).normal(0, 10, (rows, L)).astype(np.int16).ravel()
codec = d.DeltaRice(0); off = np.array([0, x.size], dtype=np.uint64)
comp, boff = codec.encode_host(x, off, M, L)
cd = torch.from_numpy(comp.copy()).cuda(); y = torch.empty(x.size, dtype=torch.int16, device="cuda"); st = torch.zeros(2, dtype=torch.int32, device="cuda")
for _ in range(6):
    codec.decode_device_async(cd, boff, off, M, L, y, st); torch.cuda.synchronize()
assert np.array_equal(y.cpu().numpy(), x)
