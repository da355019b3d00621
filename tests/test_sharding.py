"""Multi-GPU host logic on CPU: stream sharding with world_size-2 gloo processes.
No collective sits on the data path; torch.distributed only carries the barrier, a max-reduce and the host gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

from airwave_b200.sharding import gather_outputs, gather_samples, max_over_ranks, owner_of, sample_streams, stream_shard


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (4096, 8), (16384, 8), (37, 4), (5, 8), (0, 3)])
def test_shards_partition_the_streams(n, world):
    covered = []
    for r in range(world):
        first, count = stream_shard(n, world, r)
        assert count >= 0 and (not covered or covered[-1] + 1 == first or count == 0 or first == len(covered))
        covered.extend(range(first, first + count))
    assert covered == list(range(n))
    counts = [stream_shard(n, world, r)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 1
    for s in range(0, n, max(1, n // 50)):
        r = owner_of(s, n, world)
        f, c = stream_shard(n, world, r)
        assert f <= s < f + c


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, frames, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = stream_shard(n, world, rank)
        # stand-in for a per-rank engine: output of global stream g is a function of g only (streams are independent)
        g = np.arange(first, first + count, dtype=np.float32)[:, None, None]
        local = g * 10 + np.arange(2, dtype=np.float32)[None, :, None] + np.arange(frames, dtype=np.float32)[None, None, :] / 100
        dist.barrier()
        full = gather_outputs(local.astype(np.float32), n)
        slow = max_over_ranks(1.0 + rank)
        ids = sample_streams(n, world)
        rows = gather_samples(local.astype(np.float32), n, ids)
        if rank == 0:
            q.put((full, slow, ids, rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_two_rank_gloo_gather_reassembles_streams_in_order(n):
    world, frames = 2, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, slow, ids, rows = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.arange(n, dtype=np.float32)[:, None, None]
    want = g * 10 + np.arange(2, dtype=np.float32)[None, :, None] + np.arange(frames, dtype=np.float32)[None, None, :] / 100
    assert full.shape == (n, 2, frames) and np.array_equal(full, want.astype(np.float32))
    assert slow == 2.0
    # the sampled rows arrive from their owners unchanged (a negative zero or NaN payload would not survive a sum with zeros,
    # audio samples do: x + 0.0 == x bit for bit except -0.0, which compares equal)
    assert ids == sorted(set(ids)) and ids[0] == 0 and ids[-1] == n - 1
    assert np.array_equal(rows, want.astype(np.float32)[ids])
