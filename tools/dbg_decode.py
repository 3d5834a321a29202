import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import deltarice_b200 as d
codec = d.DeltaRice(0)
r = np.random.default_rng(6)
base = r.normal(0, 40, 3500*4).astype(np.int16)
for (M,L) in [(8,3500),(4,3500),(16,3500)]:
    x = torch.from_numpy(base).cuda()
    off = np.array([0, x.numel()], dtype=np.uint64)
    comp, boff = codec.encode_device(x, off, M, L)
    y = torch.zeros_like(x); st = torch.zeros(2,dtype=torch.int32,device="cuda")
    codec.decode_device_async(comp, boff, off, M, L, y, st)
    torch.cuda.synchronize()
    bad = (x!=y).nonzero().flatten()
    print("M",M,"status",int(st[0]),"mismatches",bad.numel(), "first", (int(bad[0]), int(bad[0])%L) if bad.numel() else None)
    if bad.numel():
        i=int(bad[0]); print("  x", x[i-2:i+6].tolist(), " y", y[i-2:i+6].tolist())
# locate the first bad sample per wave for M=8 and print neighbourhood of compressed words
import ctypes
M,L=8,3500
x = torch.from_numpy(base).cuda(); off = np.array([0, x.numel()], dtype=np.uint64)
comp, boff = codec.encode_device(x, off, M, L)
y = torch.zeros_like(x); st = torch.zeros(2,dtype=torch.int32,device="cuda")
codec.decode_device_async(comp, boff, off, M, L, y, st); torch.cuda.synchronize()
for w in range(4):
    b = (x[w*L:(w+1)*L]!=y[w*L:(w+1)*L]).nonzero().flatten()
    print("wave",w,"first bad", int(b[0]) if b.numel() else None, "nbad", b.numel())
# bit position of sample 1010 in wave 0
xs = base[:L].astype(np.int32); dl = np.diff(np.concatenate([[0],xs])).astype(np.int16).astype(np.int32)
u = np.where(dl>=0,2*dl,-2*dl-1); q=u>>3; ln=np.where(q>=8,25,q+4)
cs = np.concatenate([[0],np.cumsum(ln)])
for i in range(1004,1014): print(i, "len",ln[i],"bitpos",cs[i],"word",cs[i]//32,"bit",cs[i]%32, "u",u[i])
