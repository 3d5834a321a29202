import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import deltarice_b200 as d
from deltarice_b200.synth import nab_like_torch
codec = d.DeltaRice(0)
nw,L,M,wpc=153391,3500,4,2000
x = nab_like_torch(nw, L, 20251018, "cuda").reshape(-1)
off = d.chunk_offsets(wpc*L, x.numel())
comp, boff = codec.encode_device(x, off, M, L)
y = torch.empty_like(x); st = torch.zeros(2,dtype=torch.int32,device="cuda")
codec.timing(True)
for _ in range(6): codec.decode_device_async(comp, boff, off, M, L, y, st)
torch.cuda.synchronize()
t = codec.timing_read()
print({k:(round(v[0]/v[1],4) if v[1] else 0) for k,v in t.items()})
