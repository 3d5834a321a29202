"""shard.py — multi-GPU sharding of a batch of chunks (one process per GPU).

Waves and chunks are coded independently (reference src/deltaRice.c:371-373 starts every
wave from its own first sample; :193-196 a fresh bit accumulator), so a batch shards over
ranks as contiguous ranges of WHOLE chunks with no data-path exchange.  The only collective
is the one BASELINE.json's north_star names: an all-gather of each rank's compressed byte
count followed by a local exclusive scan, which gives every shard its byte offset in the
concatenated stream (chunk streams are uint32-aligned, so concatenation is pure offset
arithmetic).  torch.distributed is the plumbing: NCCL on the GPU box, gloo in CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_chunk_range(nchunks: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [c0, c1) range of whole chunks owned by `rank`."""
    base, extra = divmod(int(nchunks), int(world))
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def gather_shard_offsets(nbytes, group=None):
    """All-gather one int64 (this rank's compressed bytes) and scan.

    `nbytes`: 1-element int64 tensor on the device the process group works on (CUDA for
    NCCL, CPU for gloo); it may still be in flight on the current stream.
    Returns (counts[world], offsets[world+1]) tensors on the same device:
    shard r occupies bytes [offsets[r], offsets[r+1]) of the concatenated stream."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=nbytes.device)
    dist.all_gather_into_tensor(counts, nbytes.reshape(1).to(torch.int64), group=group)
    offsets = torch.zeros(world + 1, dtype=torch.int64, device=nbytes.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    return counts, offsets


def global_chunk_byte_offsets(local_chunk_byte_off: np.ndarray, shard_offsets: np.ndarray, rank: int) -> np.ndarray:
    """This rank's chunk byte offsets rebased into the concatenated stream."""
    return np.asarray(local_chunk_byte_off, dtype=np.uint64) + np.uint64(int(shard_offsets[rank]))


def concat_shards(streams: list[np.ndarray]) -> np.ndarray:
    """What a reader of the single stream sees: shard streams back to back (host helper used
    by tests and by callers that collect shards on one host)."""
    return np.concatenate([np.ascontiguousarray(s).view(np.uint8).ravel() for s in streams]) if streams else np.zeros(0, np.uint8)
