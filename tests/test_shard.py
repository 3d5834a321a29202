"""CPU suite, world_size 2 over gloo: the N>1 host logic — whole-chunk sharding, the byte-count
all-gather + scan, and that shard streams concatenate into the stream a single rank (the
oracle, one chunk at a time) produces.  Each rank encodes its shard with the ORACLE here (no
GPU in the CPU suite); the sharding / offset logic under test is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deltarice_b200 import shard


def test_shard_chunk_range_partitions():
    for n in (0, 1, 7, 8, 77, 1000):
        for w in (1, 2, 3, 8):
            r = [shard.shard_chunk_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    L, M, wpc, nchunks = 500, 4, 6, 7
    x = np.random.default_rng(99).normal(0, 12, nchunks * wpc * L).astype(np.int16)
    off = np.arange(nchunks + 1, dtype=np.uint64) * (wpc * L)
    c0, c1 = shard.shard_chunk_range(nchunks, world, rank)
    parts = [O.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L) for c in range(c0, c1)]
    local_boff = np.concatenate([[0], np.cumsum([4 * p.size for p in parts])]).astype(np.uint64)
    stream = np.concatenate(parts).view(np.uint8) if parts else np.zeros(0, np.uint8)
    counts, offsets = shard.gather_shard_offsets(torch.tensor([stream.size], dtype=torch.int64))
    assert int(counts[rank]) == stream.size
    assert int(offsets[0]) == 0 and int(offsets[-1]) == int(counts.sum())
    g = shard.global_chunk_byte_offsets(local_boff, offsets.numpy(), rank)
    np.save(os.path.join(tmp, f"s{rank}.npy"), stream)
    np.save(os.path.join(tmp, f"g{rank}.npy"), g)
    dist.barrier()
    if rank == 0:
        whole = shard.concat_shards([np.load(os.path.join(tmp, f"s{r}.npy")) for r in range(world)])
        want = np.concatenate([O.encode_chunk(x[int(off[c]):int(off[c + 1])], M, L) for c in range(nchunks)]).view(np.uint8)
        assert np.array_equal(whole, want)
        # every chunk is found at its global offset and starts with its sample count
        for r in range(world):
            gg = np.load(os.path.join(tmp, f"g{r}.npy"))
            a, b = shard.shard_chunk_range(nchunks, world, r)
            for i, c in enumerate(range(a, b)):
                head = whole[int(gg[i]):int(gg[i]) + 4].view(np.uint32)[0]
                assert head == wpc * L
                dec = O.decode_chunk(whole[int(gg[i]):int(gg[i + 1])].view(np.uint32), M, L)
                assert np.array_equal(dec, x[int(off[c]):int(off[c + 1])])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_concatenate(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
