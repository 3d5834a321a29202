"""Pure-torch emulation of the two host pipelines (no kernels): does chunked bidirectional traffic overlap?"""
import time, threading, torch
MB = 1 << 20
big, small, nsub = 64 * MB, 17 * MB, 16
hA = torch.empty(big * nsub, dtype=torch.uint8).pin_memory(); hB = torch.empty(big * nsub, dtype=torch.uint8).pin_memory()
hC = torch.empty(small * nsub, dtype=torch.uint8).pin_memory(); hD = torch.empty(small * nsub, dtype=torch.uint8).pin_memory()
dA = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(3)]; dB = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(3)]
dC = [torch.empty(small, dtype=torch.uint8, device="cuda") for _ in range(3)]; dD = [torch.empty(small, dtype=torch.uint8, device="cuda") for _ in range(3)]
def pipe(h_in, d_in, n_in, h_out, d_out, n_out, sync_each):
    s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    evo = [None] * 3
    for i in range(nsub):
        k = i % 3
        if evo[k] is not None: evo[k].synchronize()
        with torch.cuda.stream(s_in):
            d_in[k].copy_(h_in[i * n_in:(i + 1) * n_in], non_blocking=True)
            e = torch.cuda.Event(); e.record()
        s_k.wait_event(e)
        with torch.cuda.stream(s_k):
            d_out[k].add_(1)
            e2 = torch.cuda.Event(); e2.record()
        if sync_each: e2.synchronize()
        s_out.wait_event(e2)
        with torch.cuda.stream(s_out):
            h_out[i * n_out:(i + 1) * n_out].copy_(d_out[k], non_blocking=True)
            evo[k] = torch.cuda.Event(); evo[k].record()
    s_out.synchronize()
def enc(sync=False): pipe(hA, dA, big, hD, dD, small, sync)
def dec(sync=False): pipe(hC, dC, small, hB, dB, big, sync)
def timeit(fs, reps=3):
    for f in fs: f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        th = [threading.Thread(target=f) for f in fs]; [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("emu enc", timeit([enc]), "emu dec", timeit([dec]), "both", timeit([enc, dec]))
print("with per-sub-batch host sync: both", timeit([lambda: enc(True), lambda: dec(True)]))
# variant: process-wide copy streams (one per direction), copies enqueued only when ready
G_IN, G_OUT = torch.cuda.Stream(), torch.cuda.Stream()
def pipe2(h_in, d_in, n_in, h_out, d_out, n_out):
    s_k = torch.cuda.Stream()
    evo = [None] * 3; evk = [None] * 3
    def finish(i):
        k = i % 3
        evk[k].synchronize()
        with torch.cuda.stream(G_OUT):
            h_out[i * n_out:(i + 1) * n_out].copy_(d_out[k], non_blocking=True)
            evo[k] = torch.cuda.Event(); evo[k].record()
    for i in range(nsub + 1):
        if i < nsub:
            k = i % 3
            if evo[k] is not None: evo[k].synchronize()
            with torch.cuda.stream(G_IN):
                d_in[k].copy_(h_in[i * n_in:(i + 1) * n_in], non_blocking=True)
                e = torch.cuda.Event(); e.record()
            s_k.wait_event(e)
            with torch.cuda.stream(s_k):
                d_out[k].add_(1)
                evk[k] = torch.cuda.Event(); evk[k].record()
        if i >= 1: finish(i - 1)
    G_OUT.synchronize()
def enc2(): pipe2(hA, dA, big, hD, dD, small)
def dec2(): pipe2(hC, dC, small, hB, dB, big)
print("global copy streams: enc", timeit([enc2]), "dec", timeit([dec2]), "both", timeit([enc2, dec2]))
