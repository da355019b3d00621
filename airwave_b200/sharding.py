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


def _comm_device(group=None):
    """Where tensors must live for a collective of this group: NCCL only moves device memory, gloo moves host memory."""
    import torch
    import torch.distributed as dist

    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _equal_counts(n: int, world: int) -> bool:
    return len({stream_shard(n, world, r)[1] for r in range(world)}) == 1


def gather_outputs(local_out: np.ndarray, n_streams: int, group=None) -> Optional[np.ndarray]:
    """Host gather of per-rank output blocks [count_r][2][frames] into [n_streams][2][frames] on rank 0
    (the reference-side consumer is a host; nothing here touches the render path).  Under NCCL the blocks travel
    through device memory (NCCL rejects host tensors) and are copied back on rank 0."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = stream_shard(n_streams, world, rank)
    assert local_out.shape[0] == count
    frames = local_out.shape[2]
    dev = _comm_device(group)
    mine = torch.from_numpy(np.ascontiguousarray(local_out, np.float32)).to(dev)
    parts = [torch.empty((stream_shard(n_streams, world, r)[1], 2, frames), dtype=torch.float32, device=dev) for r in range(world)]
    if _equal_counts(n_streams, world):
        dist.all_gather(parts, mine, group=group)
    else:
        for r in range(world):
            if r == rank:
                parts[r].copy_(mine)
            dist.broadcast(parts[r], src=r, group=group)
    if rank != 0:
        return None
    return torch.cat(parts, 0).cpu().numpy()


def sample_streams(n_streams: int, world_size: int, per_rank: int = 4) -> list:
    """Global stream ids used to prove the sharding contract: for every rank the first and last streams of its
    shard and a few in between — the ones a single-GPU engine re-renders for the bit-exact comparison."""
    ids = []
    for r in range(world_size):
        first, count = stream_shard(n_streams, world_size, r)
        if count <= 0:
            continue
        k = min(per_rank, count)
        ids.extend(sorted({first + (i * (count - 1)) // max(k - 1, 1) for i in range(k)}))
    return ids


def gather_samples(local_out: np.ndarray, n_streams: int, ids, group=None) -> Optional[np.ndarray]:
    """Gathers the rows of the sampled global stream ids from their owners: [len(ids)][2][frames] on rank 0, else None.
    local_out holds this rank's shard [count][2][frames]."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = stream_shard(n_streams, world, rank)
    frames = local_out.shape[2]
    dev = _comm_device(group)
    rows = torch.zeros((len(ids), 2, frames), dtype=torch.float32, device=dev)
    for i, g in enumerate(ids):
        if first <= g < first + count:
            rows[i] = torch.from_numpy(np.ascontiguousarray(local_out[g - first], np.float32)).to(dev)
    # every row has exactly one owner and is zero elsewhere: a sum moves it without changing a value
    dist.reduce(rows, dst=0, op=dist.ReduceOp.SUM, group=group)
    return rows.cpu().numpy() if rank == 0 else None


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
