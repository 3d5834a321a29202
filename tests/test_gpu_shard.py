"""GPU suite, N > 1: ONE dataset sharded over two ranks, every rank encoding its chunk range with the
PRODUCT encoder (CUDA, through the C-ABI); the shards placed at the all-gathered offsets must be
byte-equal to the single stream the oracle produces for the whole dataset (the reference's serial
compaction, src/deltaRice.c:427-432, for every chunk in turn).

The two ranks are two processes.  With two GPUs each rank takes its own and the exchange runs over
NCCL; on a one-GPU box both ranks encode on cuda:0 and the exchange runs over gloo (NCCL refuses two
ranks on one device) — the kernels of the two ranks never wait on one another, so sharing the device
is safe.  Also here: two contexts on two devices in ONE process (launch attributes are per device)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deltarice_b200 import shard

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dataset(nchunks, wpc, L):
    from deltarice_b200.synth import nab_like
    return np.concatenate([nab_like(wpc, L, seed=4242 + c).ravel() for c in range(nchunks)])


def _rank(rank, world, port, tmp, two_gpus):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev_i = rank if two_gpus else 0
    torch.cuda.set_device(dev_i)
    dist.init_process_group("nccl" if two_gpus else "gloo", rank=rank, world_size=world)
    import deltarice_b200 as d
    L, M, wpc, nchunks = 3500, 4, 24, 9                      # 9 chunks over 2 ranks: 5 + 4
    x = _dataset(nchunks, wpc, L)
    off_all = np.arange(nchunks + 1, dtype=np.uint64) * (wpc * L)
    c0, c1 = shard.shard_chunk_range(nchunks, world, rank)
    xs = x[int(off_all[c0]):int(off_all[c1])]
    off = off_all[c0:c1 + 1] - off_all[c0]
    with d.DeltaRice(dev_i) as codec:
        comp, boff = codec.encode_host(xs, off, M, L)        # the product path (CUDA)
        assert codec.launches > 0
        nbytes = torch.tensor([comp.size], dtype=torch.int64, device=f"cuda:{dev_i}" if two_gpus else "cpu")
        counts, offsets = shard.gather_shard_offsets(nbytes)
        offsets = offsets.cpu().numpy()
        g = shard.global_chunk_byte_offsets(boff, offsets, rank)
        # the shard goes to rank 0, which places it at its gathered offset
        if rank == 0:
            whole = np.zeros(int(offsets[-1]), dtype=np.uint8)
            whole[:comp.size] = comp
            for r in range(1, world):
                t = torch.empty(int(offsets[r + 1] - offsets[r]), dtype=torch.uint8, device=nbytes.device)
                dist.recv(t, src=r)
                whole[int(offsets[r]):int(offsets[r + 1])] = t.cpu().numpy()
        else:
            dist.send(torch.from_numpy(np.ascontiguousarray(comp)).to(nbytes.device), dst=0)
        np.save(os.path.join(tmp, f"g{rank}.npy"), g)
        dist.barrier()
        if rank == 0:
            from oracle import oracle as O
            want = np.concatenate([O.encode_chunk(x[int(off_all[c]):int(off_all[c + 1])], M, L)
                                   for c in range(nchunks)]).view(np.uint8)
            assert whole.size == want.size and np.array_equal(whole, want), "concatenated product shards != oracle stream"
            gg = np.concatenate([np.load(os.path.join(tmp, f"g{r}.npy"))[:-1] for r in range(world)] + [[whole.size]]).astype(np.uint64)
            # ... and the one stream decodes, as one batch, back to the dataset
            back = codec.decode_host(whole, gg, off_all, M, L)
            assert np.array_equal(back, x)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_product_shards_equal_oracle_stream(tmp_path):
    two = torch.cuda.device_count() >= 2
    mp.spawn(_rank, args=(2, _free_port(), str(tmp_path), two), nprocs=2, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """cudaFuncSetAttribute and SM counts are per device: contexts on device 0 and 1 of one process."""
    import deltarice_b200 as d
    from oracle import oracle as O
    from deltarice_b200.synth import nab_like
    L, M = 3500, 4
    x = nab_like(40, L, seed=9).ravel()
    off = np.array([0, x.size], dtype=np.uint64)
    want = O.encode_chunk(x, M, L)
    for dev in (0, 1, 0, 1):
        with d.DeltaRice(dev) as codec:
            comp, boff = codec.encode_host(x, off, M, L)
            assert np.array_equal(comp.view(np.uint32), want)
            assert np.array_equal(codec.decode_host(comp, boff, off, M, L), x)
    # long waves and small batches take other kernels (their attributes are per device too)
    y = np.random.default_rng(3).normal(0, 9, 20000 * 3).astype(np.int16)
    offy = np.array([0, y.size], dtype=np.uint64)
    for dev in (1, 0):
        with d.DeltaRice(dev) as codec:
            comp, boff = codec.encode_host(y, offy, 8, 20000)
            assert np.array_equal(comp.view(np.uint32), O.encode_chunk(y, 8, 20000))
            assert np.array_equal(codec.decode_host(comp, boff, offy, 8, 20000), y)
