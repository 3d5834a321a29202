import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import deltarice_b200 as d
from deltarice_b200.synth import nab_like_torch
from oracle import oracle as O
codec = d.DeltaRice(0)
for (nw, L, M, wpc) in [(30000,3504,4,2000)]:
    x = nab_like_torch(nw, L, 1, "cuda").reshape(-1)
    off = d.chunk_offsets(wpc*L, x.numel())
    comp, boff = codec.encode_device(x, off, M, L)
    xh = x.cpu().numpy(); ch = comp.cpu().numpy()
    nbad=0
    for c in range(len(off)-1):
        want = O.encode_chunk(xh[int(off[c]):int(off[c+1])], M, L, mt=True)
        got = ch[int(boff[c]):int(boff[c+1])].view(np.uint32)
        if got.size!=want.size or not np.array_equal(got,want):
            nbad+=1
            if nbad==1:
                m = min(got.size,want.size); dif = np.nonzero(got[:m]!=want[:m])[0]
                print("  enc chunk",c,"sizes",got.size,want.size,"first diff word",dif[:5])
    print(nw,L,"encode bad chunks:",nbad,"of",len(off)-1)
    try:
        y = codec.decode_device(comp, boff, off, M, L)
        print("   decode", "ok" if torch.equal(x,y) else "MISMATCH")
    except Exception as e:
        print("   decode ERR", str(e)[:60])
    # host path
    try:
        ch2, boff2 = codec.encode_host(xh, off, M, L)
        print("   host encode equal dev:", np.array_equal(ch2, ch))
        y2 = codec.decode_host(ch2, boff2, off, M, L)
        print("   host decode", "ok" if np.array_equal(y2, xh) else "MISMATCH")
    except Exception as e:
        print("   host ERR", str(e)[:60])
