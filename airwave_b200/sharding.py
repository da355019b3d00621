"""Multi-GPU sharding of the binaural path: by independent stream, no collective on the data path.

GPU g of G owns the contiguous stream range [g*n/G, (g+1)*n/G); filter banks and FFT plans are replicated per
device; FDL / overlap / EQ state live only on the owner.  torch.distributed (NCCL on GPUs, gloo in the CPU tests)
is used for control only: the start barrier, the max-over-ranks of a timing, and an optional host gather of outputs.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def stream_shard(n_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the streams rank `rank` owns; ranges are contiguous, disjoint and cover [0, n_streams)."""
    if not (0 <= rank < world_size) or n_streams < 0:
        raise ValueError("bad shard request")
    first = (rank * n_streams) // world_size
    last = ((rank + 1) * n_streams) // world_size
    return first, last - first


def owner_of(stream: int, n_streams: int, world_size: int) -> int:
    """Rank that owns global stream id `stream` under stream_shard."""
    if not (0 <= stream < n_streams):
        raise ValueError("stream out of range")
    r = (stream * world_size) // n_streams
    while stream < stream_shard(n_streams, world_size, r)[0]:
        r -= 1
    while stream >= sum(stream_shard(n_streams, world_size, r)):
        r += 1
    return r


def gather_outputs(local_out: np.ndarray, n_streams: int, group=None) -> Optional[np.ndarray]:
    """Host gather of per-rank output blocks [count_r][2][frames] into [n_streams][2][frames] on rank 0
    (the reference-side consumer is a host; nothing here touches the render path)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = stream_shard(n_streams, world, rank)
    assert local_out.shape[0] == count
    frames = local_out.shape[2]
    parts = [torch.empty((stream_shard(n_streams, world, r)[1], 2, frames), dtype=torch.float32) for r in range(world)]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(local_out, np.float32)), group=group) if _equal_counts(n_streams, world) \
        else _gather_uneven(parts, local_out, group)
    if rank != 0:
        return None
    return torch.cat(parts, 0).numpy()


def _equal_counts(n: int, world: int) -> bool:
    return len({stream_shard(n, world, r)[1] for r in range(world)}) == 1


def _gather_uneven(parts, local_out, group) -> None:
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.from_numpy(np.ascontiguousarray(local_out, np.float32))
    for r in range(world):
        if r == rank:
            parts[r].copy_(mine)
        dist.broadcast(parts[r], src=r, group=group)


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
