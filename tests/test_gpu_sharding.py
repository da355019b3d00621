"""Hardware proof of the sharding contract: GPU r of G renders global streams [r*n, (r+1)*n); the per-stream
output must be bit-identical to a single-GPU render of the same global stream ids.  One process per GPU, NCCL for control
only (no collective on the data path), as bench.py --gpus N runs it.  Skipped when the box has fewer than two GPUs."""
import os
import socket
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import airwave_b200 as aw
    import bench
    rank_, world_, local, dist = bench.init_ranks(torch)
    try:
        pcm, rate = bench.hrir_pcm("RoomSH1.0")
        l_idx, r_idx = bench.speaker_maps(8)
        out = []
        for block, n in ((256, 300), (64, 77), (1024, 40)):        # full tiles, cut tiles, large blocks
            res = bench.shard_check(aw, torch, dist, rank_, world_, local, n, 8, block, (pcm, rate, bench.FS, l_idx, r_idx, block), 4 * block)
            out.append((block, n, res))
        if rank == 0:
            q.put(out)
    finally:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()


def _run(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return out


def _device_count():
    import airwave_b200 as aw
    return aw.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_streams_rendered_on_gpu_r_of_g_are_bit_identical_to_a_single_gpu_render(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    for block, n, res in _run(world):
        assert res["global_streams"] == world * n and len(res["per_rank_checksum"]) == world
        assert len(res["sampled_ids"]) >= 2 * world
        assert res["bit_identical_to_single_gpu"], (block, n, res)
        assert len(set(res["per_rank_checksum"])) == world, "every rank must have rendered its own streams"


def test_single_gpu_shard_check_runs():
    """The same check degenerates gracefully on one GPU (the sampled ids are re-rendered by a second engine)."""
    import torch
    import airwave_b200 as aw
    import bench
    pcm, rate = bench.hrir_pcm("RoomSH1.0")
    l_idx, r_idx = bench.speaker_maps(8)
    res = bench.shard_check(aw, torch, None, 0, 1, 0, 300, 8, 256, (pcm, rate, bench.FS, l_idx, r_idx, 256), 1024)
    assert res["bit_identical_to_single_gpu"] and res["sampled_ids"][0] == 0 and res["sampled_ids"][-1] == 299
