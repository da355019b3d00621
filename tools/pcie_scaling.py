"""Concurrent host<->device copy floor on N GPUs of one box, for the transfer sizes of the end-to-end step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_scaling.py

Every rank owns one GPU and moves, per step, what bench.py's e2e step moves per GPU (C2: 134 MB host->device, 34 MB
device->host, pinned host memory), with nothing else running: the best any engine could do on this platform.  Variants:
input buffers page-locked (cudaHostAllocDefault) or write-combined (cudaHostAllocWriteCombined); device->host alone (the
offline render of BASELINE.json configs[4] only returns output).  Rank 0 prints one line per variant: per-rank GB/s
(min/median/max) and the aggregate — and the stream-s/s the C2 wire format would reach at that rate.
"""
import ctypes as C
import os
import statistics
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(name); break
    except OSError:
        pass
if rt is None:
    import glob
    rt = C.CDLL(sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")))[0])
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaFreeHost.argtypes = [C.c_void_p]
H2D, D2H = 1, 2
N_IN, N_OUT = 4096 * 8 * 1024 * 4, 4096 * 2 * 1024 * 4          # bytes per e2e step and GPU (C2, 1024-frame submits)


def host(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), nbytes, flags) == 0
    C.memset(p, 1, nbytes)
    return p


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(name, h2d_bytes, d2h_bytes, flags, iters=24):
    hin = [host(h2d_bytes, flags) for _ in range(2)] if h2d_bytes else []
    hout = [host(d2h_bytes, 0) for _ in range(2)] if d2h_bytes else []
    din = [torch.empty(max(h2d_bytes, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    dout = [torch.ones(max(d2h_bytes, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def loop(n):
        for i in range(n):
            if h2d_bytes:
                assert rt.cudaMemcpyAsync(din[i % 2].data_ptr(), hin[i % 2], h2d_bytes, H2D, s1.cuda_stream) == 0
            if d2h_bytes:
                assert rt.cudaMemcpyAsync(hout[i % 2], dout[i % 2].data_ptr(), d2h_bytes, D2H, s2.cuda_stream) == 0
        torch.cuda.synchronize()

    loop(4)
    barrier()
    t0 = time.perf_counter()
    loop(iters)
    dt = (time.perf_counter() - t0) / iters
    barrier()
    rec = torch.tensor([dt], dtype=torch.float64)
    allr = [torch.zeros_like(rec) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, rec)
    else:
        allr = [rec]
    ts = [float(t.item()) for t in allr]
    if rank == 0:
        worst = max(ts)
        gin = [h2d_bytes / t / 1e9 for t in ts]; gout = [d2h_bytes / t / 1e9 for t in ts]
        per_gpu_streams = 4096 * (1024 / 48000.0) / worst if h2d_bytes else (d2h_bytes / (2 * 4 * 48000.0)) / worst
        print(f"N={world} {name:34s} step {worst * 1e3:7.3f} ms (slowest rank)  H2D/rank GB/s min {min(gin):5.1f} med {statistics.median(gin):5.1f} max {max(gin):5.1f}"
              f"  D2H/rank min {min(gout):5.1f} med {statistics.median(gout):5.1f} max {max(gout):5.1f}  aggregate in+out {world * (h2d_bytes + d2h_bytes) / worst / 1e9:6.1f} GB/s"
              f"  => floor {world * per_gpu_streams:9.0f} stream-s/s", flush=True)
    for p in hin + hout:
        rt.cudaFreeHost(p)


run("C2 e2e step: H2D 134 MB + D2H 34 MB", N_IN, N_OUT, 0)
run("same, write-combined input buffers", N_IN, N_OUT, 4)
run("H2D 134 MB alone", N_IN, 0, 0)
run("D2H 67 MB alone (offline render)", 0, 2 * N_OUT, 0)
if world > 1:
    dist.destroy_process_group()
