#!/usr/bin/env python
"""bench.py — binaural stream-seconds rendered per second (7.1 -> 2 ch, 48 kHz) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

A step = one block of B frames rendered for every stream of the workload (one pass of the hot
path: K2 input_rfft -> K3 fdl_cmac -> K4 irfft_out).  The default workload is BASELINE.json
configs[1] ("C2"): 7.1 (8 virtual speakers) -> binaural, RoomSH1.0 HeSuVi 14-ch HRIR, 256-frame
blocks, 4,096 concurrent streams per GPU (weak scaling: streams are independent, no collective).

`value`  : whole-job stream-s/s with inputs resident in HBM (device events, max over ranks).
`e2e`    : same metric through aw_engine_submit/aw_engine_wait with pinned HOST buffers, every
           step's input copied H2D and its output copied D2H inside the timed region.
`roofline`: the dominant kernel (the block kernel k_persistent; on the three-kernel path the slowest of
           K2/K3/K4): algorithmic bytes per launch / its mean duration (CUDA events around each
           launch on the engine's stream) vs MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (C restatement of the reference algorithm) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000.0
SEED = 0x41495257
METRIC = "binaural stream-sec/sec (7.1->2ch, 48 kHz)"
UNIT = "stream-s/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    # name: (description, speakers, block, streams per GPU, hrir)
    "C2": ("7.1->binaural, RoomSH1.0 HeSuVi 14-ch HRIR (4320 taps), B=256, 4096 streams/GPU", 8, 256, 4096, "RoomSH1.0"),
    "C1": ("stereo->binaural, NeutralSH1.0, B=512, single stream", 2, 512, 1, "NeutralSH1.0"),
    "C3": ("7.1->binaural, synthetic 65,536-tap BRIR, B=512, 1024 streams/GPU", 8, 512, 1024, "synthetic65536"),
    "C4": ("full chain: 7.1->binaural with StageSH1.0 relabelled 44.1 kHz and resampled to 48 kHz (4320 -> 4702 taps), B=256, "
           "+ 10-band parametric EQ (CCA CRA fixture), 8192 streams/GPU", 8, 256, 8192, "StageSH1.0@44100"),
    "F3": ("per-device profiles at batch scale (SURVEY.md 8(f3)): 4096 streams in 64 ranges, each bound to one of two HRIR banks "
           "(RoomSH1.0 / StageSH1.0) and one of two equalizers (CCA CRA fixture / Bass Booster), 7.1->binaural, B=256",
           8, 256, 4096, "RoomSH1.0+StageSH1.0"),
    "C5-64": ("7.1->binaural, RoomSH1.0, B=64, 2048 streams/GPU", 8, 64, 2048, "RoomSH1.0"),
    "C5-128": ("7.1->binaural, RoomSH1.0, B=128, 2048 streams/GPU", 8, 128, 2048, "RoomSH1.0"),
    "C5-512": ("7.1->binaural, RoomSH1.0, B=512, 2048 streams/GPU", 8, 512, 2048, "RoomSH1.0"),
    "C5-1024": ("7.1->binaural, RoomSH1.0, B=1024, 2048 streams/GPU", 8, 1024, 2048, "RoomSH1.0"),
    "C5-2048": ("7.1->binaural, RoomSH1.0, B=2048, 2048 streams/GPU", 8, 2048, 2048, "RoomSH1.0"),
    "C5-4096": ("7.1->binaural, RoomSH1.0, B=4096, 2048 streams/GPU", 8, 4096, 2048, "RoomSH1.0"),
    # BASELINE.json configs[4] as specified: offline batch render, 2,048 streams per GPU x 60 s, block-size sweep; input synthesised
    # on the device chunk by chunk (188.7 GB per GPU would not cross the host link), output copied to
    # pinned host memory through aw_engine_submit_device / aw_engine_wait.  The block size in this tuple is the headline entry.
    "C5-offline": ("offline batch render, 7.1->binaural, RoomSH1.0, 2048 streams/GPU x 60 s, block sweep 64..4096, device-synthesised "
                   "input, output to pinned host memory", 8, 256, 2048, "RoomSH1.0"),
}


def hrir_pcm(name: str):
    """Returns (pcm [channels][frames] float32, sample_rate)."""
    import numpy as np
    if name == "synthetic65536":   # 0.5*delta[n-190] + 0.05*N(0,1)*exp(-n/(0.25 fs)), seed 1
        rng = np.random.default_rng(1)
        n = np.arange(65536)
        pcm = (0.05 * rng.standard_normal((14, 65536)) * np.exp(-n / (0.25 * FS))).astype(np.float32)
        pcm[:, 190] += 0.5
        return pcm, FS
    import airwave_b200 as aw
    name, _, relabel = name.partition("@")      # "preset@rate": treat the file as a source of that sample rate
    wav = aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", name + ".wav"))
    return wav.audioData, (float(relabel) if relabel else wav.sampleRate)


def eq_definition_for(workload: str, parser):
    """C4 carries the reference's 10-filter fixture (AirwaveTests/Fixtures/CCA CRA ParametricEq.txt); others have no EQ."""
    if workload != "C4":
        return None
    with open(os.path.join(GOLDEN, "eq", "CCA CRA ParametricEq.txt"), "rb") as f:
        return parser(f.read(), "CCA CRA ParametricEq.txt")


def recorded_traffic(workload: str, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/), or None."""
    best = None
    try:
        for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
            if name.endswith("_traffic.json"):
                rec = json.load(open(os.path.join(ROOT, "profiles", name))).get(workload)
                if rec and rec.get("kernel") == kernel:
                    best = float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])
    except Exception:
        return None
    return best


def speaker_maps(S: int):
    import airwave_b200 as aw
    lay = aw.InputLayout.detect(S)
    m = aw.HRIRChannelMap.hesuvi14Channel(lay.channels)
    return [m.getIndices(s)[0] for s in lay.channels], [m.getIndices(s)[1] for s in lay.channels]


def algorithmic_bytes(S: int, B: int, P: int, eq_filters: int = 0) -> int:
    """per stream per block, FDL ring (1 slot written + P-1 read) + input + output [+ EQ state: 4 doubles per
    filter read and written, which is what the table's 322,176 B for C4 contains]."""
    return 8 * S * B * P + 4 * S * B + 8 * B + 64 * eq_filters


def call_minimum_bytes(S: int, B: int, P: int, k: int) -> int:
    """HBM bytes a k-block call cannot avoid, per stream: the P-1 history slots read once, k new slots written, k blocks of input
    read, k blocks of output written (with k = 1 this is algorithmic_bytes)."""
    return 8 * S * B * (P - 1) + k * (8 * S * B + 4 * S * B + 8 * B)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampled every 100 ms from before the warm-up (so that it is already running when a short timed region starts);
    stop(t0, t1) keeps the samples whose timestamps fall inside the timed region."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def lines(self) -> int:
        try:
            return sum(1 for _ in open(self.file.name))
        except Exception:
            return 0

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.strip().split(",") for r in open(self.file.name).read().splitlines() if r.strip()]
        os.unlink(self.file.name)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside, every = [], []
        for r in rows:
            try:
                stamp = time.mktime(time.strptime(r[0].strip().split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + r[0].strip().split(".")[1])
                rec = (float(r[2]), float(r[3]), [n for n, v in zip(names, r[6:10]) if v.strip().lower().startswith("active")])
            except Exception:
                continue
            every.append(rec)
            if t0 - 0.1 <= stamp <= t1 + 0.1:
                inside.append(rec)
        # a timed region shorter than the sampling period may catch no sample: fall back to the samples taken under the same load
        # during the warm-up that precedes it, and say so
        used, window = (inside, "timed region") if inside else (every[-3:], "same load just before/after the timed region (shorter than the sampling period)")
        sm = [u[0] for u in used]
        reasons = sorted({n for u in used for n in u[2]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(u[1] for u in used) if used else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def pin_to_gpu_numa_node(gpu_index: int):
    """Binds this rank to the CPU cores nvidia-smi reports as local to its GPU (pinned host buffers then live on that NUMA node:
    the end-to-end path is PCIe/host-memory bound).  Returns a description, or None when the topology cannot be read."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        header = next(l for l in out.splitlines() if "CPU Affinity" in l)
        cols = [c.strip() for c in header.split("\t")]
        col = cols.index("CPU Affinity")
        row = next(l for l in out.splitlines() if l.startswith(f"GPU{gpu_index}\t") or l.startswith(f"GPU{gpu_index} "))
        spec = [c.strip() for c in row.split("\t")][col]
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


def cpu_filters(S, pcm, rate, l_idx, r_idx):
    """[S][2][taps] impulse responses the reference would build its engines from (resampled like Resampler.swift when rates differ)."""
    import numpy as np
    import oracle
    chans = {}
    for c in set(l_idx) | set(r_idx):
        chans[c] = pcm[c] if abs(rate - FS) < 0.01 else oracle.resample_high_quality(pcm[c], rate, FS)
    taps = len(next(iter(chans.values())))
    h = np.zeros((S, 2, taps), np.float32)
    for s in range(S):
        h[s, 0], h[s, 1] = chans[l_idx[s]], chans[r_idx[s]]
    return h


def cpu_eq_definition(workload: str):
    """The reference arm's equalizer for the full-chain workload (parsed by the oracle's parser: no CUDA library involved)."""
    if workload != "C4":
        return None
    import oracle
    with open(os.path.join(GOLDEN, "eq", "CCA CRA ParametricEq.txt"), "rb") as f:
        return oracle.parse_equalizer_apo(f.read(), "CCA CRA ParametricEq.txt")


def cpu_sample(S, B, pcm, rate, l_idx, r_idx, budget_s, threads=0, eq_definition=None):
    """Times the oracle (C restatement of the reference) on a bounded sample of the workload."""
    import numpy as np
    import oracle
    h = cpu_filters(S, pcm, rate, l_idx, r_idx)
    cores = oracle.max_threads() if threads <= 0 else threads
    streams = max(cores * 4, 8)
    batch = oracle.CpuBatch(streams, S, B, h)
    if eq_definition is not None:
        batch.set_eq(eq_definition, FS)
    sec, _ = batch.step(2, cores)                          # pilot (also warms the FDL)
    per_block = max(sec / 2, 1e-6)
    blocks = int(max(4, min(200000, budget_s / per_block)))     # ~budget_s seconds of CPU work
    sec, _ = batch.step(blocks, cores)
    value = streams * blocks * (B / FS) / sec
    return value, cores, f"{streams} streams x {blocks} blocks of {B} frames, one stream per thread, {cores} threads, {sec:.2f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import oracle
    desc, S, B, n_per_gpu, hrir = WORKLOADS[args.workload]
    pcm, rate = hrir_pcm_cpu(hrir.split("+")[0])
    l_idx, r_idx = hesuvi14_cpu(S)
    h = cpu_filters(S, pcm, rate, l_idx, r_idx)
    taps = h.shape[2]
    P = -(-taps // B)
    cores = oracle.max_threads()
    streams = max(cores * 4, 8)
    eq_def = cpu_eq_definition(args.workload)
    batch = oracle.CpuBatch(streams, S, B, h)
    if eq_def is not None:
        batch.set_eq(eq_def, FS)
    # size the per-step sample: long enough (~0.25 s) that thread start-up does not count against the CPU, short enough
    # that the whole --steps/--warmup run ends within a few minutes even for long BRIRs
    pilot, _ = batch.step(8, cores)
    pilot, _ = batch.step(8, cores)
    per_block = max(pilot / 8, 1e-6)
    step_budget = min(0.25, 150.0 / max(args.steps + args.warmup, 1))
    blocks_per_step = int(max(1, min(4000, step_budget / per_block)))
    while per_block * blocks_per_step * (args.steps + args.warmup) > 150 and streams > cores:
        streams //= 2
        batch = oracle.CpuBatch(streams, S, B, h)
        if eq_def is not None:
            batch.set_eq(eq_def, FS)
        pilot, _ = batch.step(8, cores)
        per_block = max(pilot / 8, 1e-6)
    for _ in range(args.warmup):
        batch.step(blocks_per_step, cores)
    t = 0.0
    for _ in range(args.steps):
        sec, _ = batch.step(blocks_per_step, cores)
        t += sec
    value = streams * blocks_per_step * args.steps * (B / FS) / t
    sample = f"{streams} streams x {blocks_per_step} blocks of {B} frames per step, one stream per thread, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "speakers": S, "block": B, "partitions": P, "taps": taps,
                   "note": "reference algorithm (one ConvolutionEngine per speaker x ear, zvmul+zvadd passes) restated in C "
                           "(oracle/airwave_oracle.c); the Swift/vDSP reference cannot run on Linux"
                           + ("; full chain: every stream's output goes through its own 10-biquad Double cascade" if eq_def is not None else "")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def hrir_pcm_cpu(name: str):
    """hrir_pcm without touching the CUDA library (reference arm)."""
    import numpy as np
    import oracle
    if name == "synthetic65536":
        rng = np.random.default_rng(1)
        n = np.arange(65536)
        pcm = (0.05 * rng.standard_normal((14, 65536)) * np.exp(-n / (0.25 * FS))).astype(np.float32)
        pcm[:, 190] += 0.5
        return pcm, FS
    name, _, relabel = name.partition("@")
    w = oracle.load_wav(os.path.join(GOLDEN, "hrtf", name + ".wav"))
    return w.audioData, (float(relabel) if relabel else w.sampleRate)


def hesuvi14_cpu(S: int):
    import oracle
    lay = oracle.InputLayout.detect(S)
    m = oracle.HRIRChannelMap.hesuvi14Channel(lay.channels)
    return [m.getIndices(s)[0] for s in lay.channels], [m.getIndices(s)[1] for s in lay.channels]


def percentile(sorted_values, q):
    return sorted_values[min(len(sorted_values) - 1, int(q * len(sorted_values)))]


def init_ranks(torch):
    """(rank, world, local, dist-or-None) from the torchrun environment; one process per GPU, NCCL for control only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def shard_check(aw, torch, dist, rank, world, local, n, S, B, bank_args, frames):
    """Hardware proof of the sharding contract: rank r renders global streams [r*n, (r+1)*n) on GPU r; a
    single-GPU engine on rank 0 re-renders a sample of the same global ids; the outputs must be bit-identical."""
    import numpy as np
    from airwave_b200 import sharding
    bank = aw.HRIRBank(*bank_args, device=local)
    eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=frames, max_partitions=bank.partitions, device=local)
    eng.set_bank(bank)
    x = torch.empty((n, S, frames), dtype=torch.float32, device=f"cuda:{local}")
    y = torch.empty((n, 2, frames), dtype=torch.float32, device=f"cuda:{local}")
    aw._lib.check(aw.lib().aw_synth_fill_device(local, x.data_ptr(), rank * n, n, S, 0, frames, SEED, None))
    torch.cuda.synchronize()
    eng.process_device(x.data_ptr(), S * frames, frames, y.data_ptr(), 2 * frames, frames, frames)
    torch.cuda.synchronize()
    mine = y.cpu().numpy()
    eng.close()
    total = world * n
    ids = sharding.sample_streams(total, world, per_rank=4)
    checks = torch.tensor([float(mine.astype("float64").sum())], dtype=torch.float64, device=f"cuda:{local}")
    all_checks = [torch.zeros_like(checks) for _ in range(world)]
    if dist is not None:
        rows = sharding.gather_samples(mine, total, ids)
        dist.all_gather(all_checks, checks)
    else:
        rows, all_checks = mine[ids], [checks]
    result = None
    if rank == 0:
        ref = aw.BinauralEngine(len(ids), S, B, FS, max_frames_per_call=frames, max_partitions=bank.partitions, device=local)
        ref.set_bank(bank)
        xr = torch.empty((len(ids), S, frames), dtype=torch.float32, device=f"cuda:{local}")
        yr = torch.empty((len(ids), 2, frames), dtype=torch.float32, device=f"cuda:{local}")
        for i, g in enumerate(ids):          # stream g's signal depends on its GLOBAL id only
            aw._lib.check(aw.lib().aw_synth_fill_device(local, xr[i].data_ptr(), g, 1, S, 0, frames, SEED, None))
        torch.cuda.synchronize()
        ref.process_device(xr.data_ptr(), S * frames, frames, yr.data_ptr(), 2 * frames, frames, frames)
        torch.cuda.synchronize()
        want = yr.cpu().numpy()
        ref.close()
        result = {"global_streams": total, "sampled_ids": ids, "frames": frames,
                  "bit_identical_to_single_gpu": bool(np.array_equal(rows, want)) and bool(np.abs(want).max() > 0),
                  "per_rank_checksum": [float(c.item()) for c in all_checks]}
    return result


def run_sync_latency(aw, local):
    """The reference's real-time contract is synchronous: process(inputLeft:...frameCount:) returns filled buffers
    (AudioPipeline.swift:3-11), one 512-frame block per callback = 10.7 ms of audio (HRIRManager.swift:149).  Host->host latency
    of aw_engine_process_stereo (C1) and aw_engine_process (n streams of 7.1, B = 256), with and without zero-copy staging."""
    import ctypes as C
    import numpy as np
    out = {}
    lib = aw.lib()
    wav = aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", "NeutralSH1.0.wav"))

    def timed(call, reps):
        for _ in range(200):
            call()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter_ns()
            call()
            t.append((time.perf_counter_ns() - t0) * 1e-3)
        t.sort()
        return {"p50_us": round(percentile(t, 0.5), 2), "p99_us": round(percentile(t, 0.99), 2), "mean_us": round(sum(t) / len(t), 2), "calls": reps}

    for mode in ("zero_copy", "staged_copies"):
        os.environ["AW_ZERO_COPY"] = "1" if mode == "zero_copy" else "0"
        res = {}
        bank = aw.HRIRBank.from_wav(wav, FS, aw.InputLayout.stereo(), 512, device=local)
        for frames in (512, 2048):
            eng = aw.BinauralEngine(1, 2, 512, FS, max_frames_per_call=4096, max_partitions=bank.partitions, device=local, literal_stereo=True)
            eng.set_bank(bank)
            l = (np.random.default_rng(1).random(frames, dtype=np.float32) - 0.5) * 0.5
            r = (np.random.default_rng(2).random(frames, dtype=np.float32) - 0.5) * 0.5
            ol, orr = np.zeros(frames, np.float32), np.zeros(frames, np.float32)
            args = (eng._h, l.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p),
                    orr.ctypes.data_as(C.c_void_p), frames)
            res[f"process_stereo_{frames}_frames"] = timed(lambda: lib.aw_engine_process_stereo(*args), 10000 if frames == 512 else 3000)
            eng.close()
        wav71 = aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", "RoomSH1.0.wav"))
        bank71 = aw.HRIRBank.from_wav(wav71, FS, aw.InputLayout.surround71(), 256, device=local)
        for n in (1, 16, 64):
            eng = aw.BinauralEngine(n, 8, 256, FS, max_frames_per_call=256, max_partitions=bank71.partitions, device=local)
            eng.set_bank(bank71)
            x = ((np.random.default_rng(3).random((n, 8, 256), dtype=np.float32) - 0.5) * 0.5)
            y = np.zeros((n, 2, 256), np.float32)
            args = (eng._h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 256)
            res[f"process_{n}_streams_7.1_256_frames"] = timed(lambda: lib.aw_engine_process(*args), 3000)
            eng.close()
        out[mode] = res
    os.environ.pop("AW_ZERO_COPY", None)
    p50 = out["zero_copy"]["process_stereo_512_frames"]["p50_us"]
    out["callback_period_us"] = 512 / FS * 1e6
    out["headroom_vs_10.7ms_callback"] = round(512 / FS * 1e6 / p50, 1)
    out["how"] = ("time.perf_counter_ns around the ctypes call (host pointers in ordinary pageable memory, result in the caller's buffer on "
                  "return); zero_copy = kernels read/write page-locked mapped staging directly, staged_copies = cudaMemcpyAsync both ways")
    return out


def run_offline(args):
    """BASELINE.json configs[4]: offline batch render of 2,048 streams per GPU x `--offline-seconds` s, one pass per block size."""
    import numpy as np
    import torch
    import airwave_b200 as aw
    from airwave_b200 import sharding

    if not torch.cuda.is_available() or aw.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    rank, world, local, dist = init_ranks(torch)
    affinity = pin_to_gpu_numa_node(local) if world > 1 else None
    desc, S, B_head, n, hrir = WORKLOADS[args.workload]
    if args.streams > 0:
        n = args.streams
    pcm, rate = hrir_pcm(hrir)
    l_idx, r_idx = speaker_maps(S)
    blocks = [int(b) for b in args.offline_blocks.split(",")]
    seconds = args.offline_seconds
    F = 4096                                           # frames per call (chunk): the adapter's maximum (RealtimeAudioProcessor.swift:85)
    chunks = max(2, int(round(seconds * FS / F)))
    rendered_s = chunks * F / FS
    x = torch.empty((n, S, F), dtype=torch.float32, device=f"cuda:{local}")
    hout = [aw.PinnedBuffer((n, 2, F)) for _ in range(2)]
    sampler = ClockSampler(local)
    sampler.start()
    t_all0 = time.time()
    sweep = []
    launches = 0
    for B in blocks:
        bank = aw.HRIRBank(pcm, rate, FS, l_idx, r_idx, B, device=local)
        eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=F, max_partitions=bank.partitions, device=local, pipelined=True)
        eng.set_bank(bank)
        st = eng.cuda_stream
        xp = x.data_ptr()

        def chunk(c):
            aw._lib.check(aw.lib().aw_synth_fill_device(local, xp, rank * n, n, S, c * F, F, SEED, st))
            eng.submit_device(xp, S * F, F, hout[c % 2].array.ctypes.data, F)

        for c in range(3):                             # warm-up: fills the FDL (P*B <= 4352 frames) and the pipeline
            chunk(c)
        eng.wait()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        c0 = eng.counters()
        t0 = time.perf_counter()
        for c in range(chunks):
            chunk(3 + c)
        eng.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        c1 = eng.counters()
        launches += int(c1["kernel_launches"] - c0["kernel_launches"])
        last = hout[(3 + chunks - 1) % 2].array
        checksum = float(last[:, :, -1].astype("float64").sum())
        dt_max = sharding.max_over_ranks(dt) if dist is not None else dt
        # the whole job's output is reassembled on the host in global stream order ("the host gathers each GPU's
        # output"); outside the timed region, last 64 frames of the render
        gathered = sharding.gather_outputs(np.ascontiguousarray(last[:, :, -64:]), world * n) if dist is not None else last[:, :, -64:]
        gathered_shape = list(gathered.shape) if gathered is not None else None
        gathered_checksum = float(gathered.astype("float64").sum()) if gathered is not None else None
        # device time per call (events around 40 calls, nothing else on the stream)
        stream = torch.cuda.ExternalStream(st, device=local)
        y = torch.empty((n, 2, F), dtype=torch.float32, device=f"cuda:{local}")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
        ev[0].record(stream)
        for j in range(40):
            eng.process_device(xp, S * F, F, y.data_ptr(), 2 * F, F, F)
            ev[j + 1].record(stream)
        torch.cuda.synchronize()
        per_call = sorted(ev[j].elapsed_time(ev[j + 1]) for j in range(40))
        nb = F // B
        peak, _ = measured_peaks()
        dev_ms = per_call[len(per_call) // 2]
        entry = {"block": B, "partitions": bank.partitions, "block_period_ms": 1e3 * B / FS,
                 "e2e_value": world * n * rendered_s / dt_max, "elapsed_s": dt_max, "rendered_s_per_stream": rendered_s,
                 "device_ms_per_call": {"p50": dev_ms, "p99": percentile(per_call, 0.99)},
                 "device_ms_per_block": {"p50": dev_ms / nb, "p99": percentile(per_call, 0.99) / nb},
                 "device_value_per_gpu": n * (F / FS) / (dev_ms * 1e-3),
                 "roofline_frac": nb * n * algorithmic_bytes(S, B, bank.partitions) / (dev_ms * 1e-3) / 1e9 / peak,
                 "d2h_gbs_per_gpu": n * 2 * F * 4 * chunks / dt_max / 1e9, "kernels": eng.plan()["kernels"], "checksum_rank0": checksum,
                 "host_gather": {"shape": gathered_shape, "checksum": gathered_checksum}}
        sweep.append(entry)
        eng.close()
        del bank
    t_all1 = time.time()
    clocks = sampler.stop(t_all0, t_all1)
    if rank == 0:
        head = next((e for e in sweep if e["block"] == B_head), sweep[0])
        peak, peak_src = measured_peaks()
        line = {
            "metric": METRIC, "value": head["device_value_per_gpu"] * world, "unit": UNIT, "n_gpus": world, "steps": 40, "warmup": 3,
            "ms_per_step": head["device_ms_per_call"]["p50"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (generated on the device, counter-based, keyed by global stream id)",
            "config": {"workload": f"{args.workload}: {desc}", "streams_per_gpu": n, "speakers": S, "headline_block": head["block"],
                       "seconds_per_stream": rendered_s, "frames_per_call": F,
                       "step": f"one {F}-frame call ({F // head['block']} blocks, one launch) for all streams + its output copied to the host",
                       "l2": "inputs larger than L2: FDL working set 571 MB+ per GPU"},
            "e2e": {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * 2 * F * 4,
                    "timing": "host wall clock around the render loop (device synth -> aw_engine_submit_device -> pinned host), "
                              "synchronised on both sides, max over ranks"},
            "roofline": {"bound": "hbm", "kernel": head["kernels"][0], "achieved": head["roofline_frac"] * peak, "peak": peak, "unit": "GB/s",
                         "frac": head["roofline_frac"], "traffic": None, "peak_source": peak_src,
                         "note": "algorithmic bytes of SURVEY.md 8(d) per block; a 4096-frame call re-reads a tile's FDL rows from L2, "
                                 "so the fraction may exceed what one block per launch can reach"},
            "sweep": sweep, "gpu_launches": launches, "host_affinity": affinity, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import airwave_b200 as aw

    if not torch.cuda.is_available() or aw.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    rank, world, local, dist = init_ranks(torch)
    affinity = pin_to_gpu_numa_node(local) if world > 1 else None

    desc, S, B, n, hrir = WORKLOADS[args.workload]
    if args.streams > 0:
        n = args.streams
    presets = hrir.split("+")
    pcm, rate = hrir_pcm(presets[0])
    l_idx, r_idx = speaker_maps(S)
    bank = aw.HRIRBank(pcm, rate, FS, l_idx, r_idx, B, device=local)
    banks = [bank] + [aw.HRIRBank(*hrir_pcm(name), FS, l_idx, r_idx, B, device=local) for name in presets[1:]]
    P, taps = bank.partitions, bank.taps
    e2e_frames = max(B, min(4096, args.e2e_frames // B * B))
    # workloads with an equalizer: its float64 cascade runs on an internal stream next to the next call's convolution
    # (AW_ENGINE_OVERLAP_EQ; outputs alternate between two buffers, as that mode requires)
    overlap_eq = args.workload in ("C4", "F3") and not args.serial_eq
    eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=e2e_frames, max_partitions=P, device=local, pipelined=True,
                            overlap_eq=overlap_eq)
    if len(banks) == 1:
        eng.set_bank(bank)
    else:                                     # 64 ranges, banks alternating: one grid per range, run concurrently on side streams
        ranges = 64
        per = n // ranges
        for r in range(ranges):
            eng.set_bank(banks[r % len(banks)], r * per, per if r < ranges - 1 else n - r * per)
    eq_def = eq_definition_for(args.workload, aw.EqualizerAPOParser.parse)
    eq_filters = 0
    if eq_def is not None:                    # full chain: spatial -> EQ (AudioEffectGraph.swift:195-210)
        eng.eq_prepare(eq_def)
        eq_filters = sum(1 for f in eq_def["filters"] if f.get("isEnabled", True))
    if args.workload == "F3":                 # a profile = (HRIR preset, EQ preset): every range also gets one of two equalizers
        defs = []
        for name in ("CCA CRA ParametricEq.txt", "Bass Booster.txt"):
            with open(os.path.join(GOLDEN, "eq", name), "rb") as f:
                defs.append(aw.EqualizerAPOParser.parse(f.read(), name))
        ranges = 64
        per = n // ranges
        for r in range(ranges):
            eng.eq_prepare(defs[(r // 2) % 2], r * per, per if r < ranges - 1 else n - r * per)
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=local)

    # device-resident synthetic input: a time-contiguous ring of R blocks per (stream, speaker); a step is one call of kb blocks
    kb = e2e_frames // B if args.blocks_per_call <= 0 else max(1, min(args.blocks_per_call, e2e_frames // B))
    R = max(8, 2 * kb)
    x = torch.empty((n, S, R * B), dtype=torch.float32, device=f"cuda:{local}")
    y = torch.empty((2, n, 2, kb * B), dtype=torch.float32, device=f"cuda:{local}")
    aw._lib.check(aw.lib().aw_synth_fill_device(local, x.data_ptr(), rank * n, n, S, 0, R * B, SEED, None))
    torch.cuda.synchronize()
    xp, yp = x.data_ptr(), y.data_ptr()
    y_alt = 4 * n * 2 * kb * B if overlap_eq else 0          # byte offset of the second output buffer

    def step(j: int):
        eng.process_device(xp + 4 * (j % (R // kb)) * kb * B, S * R * B, R * B, yp + (j & 1) * y_alt, 2 * kb * B, kb * B, kb * B)

    W = max(args.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()
    for j in range(W):
        step(j)
    torch.cuda.synchronize()

    K = args.steps
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for j in range(W):                       # keep the load on while nvidia-smi comes up (short timed regions)
        step(j)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    c0 = eng.counters()
    t_start = time.time()
    ev_start.record(stream)
    for j in range(K):                       # exactly K steps back to back: nothing but the engine's launches between the two events
        step(W + j)
    if overlap_eq:
        eng.flush()                          # the last call's equalizer is part of the timed region
    ev_end.record(stream)
    torch.cuda.synchronize()
    t_end = time.time()
    c1 = eng.counters()
    if dist is not None:
        dist.barrier()
    if sampler.proc is not None and sampler.lines() < 2:
        # nvidia-smi had not produced a sample yet (it takes a few hundred ms to come up on an 8-GPU box): keep the very same
        # load running, untimed, until it has — at most 2 s
        t_wait = time.time()
        while sampler.lines() < 2 and time.time() - t_wait < 2.0:
            for j in range(20):
                step(j)
            torch.cuda.synchronize()
    clocks = sampler.stop(t_start, t_end)
    elapsed_ms = ev_start.elapsed_time(ev_end)
    # per-block latency distribution: a separate pass with an event after every step (the events would otherwise sit between
    # the launches of the timed region)
    KL = min(max(K, 200), 1000)              # enough samples for a p99 even when the driver asks for 20 steps
    events = [torch.cuda.Event(enable_timing=True) for _ in range(KL + 1)]
    events[0].record(stream)
    for j in range(KL):
        step(W + K + j)
        events[j + 1].record(stream)
    torch.cuda.synchronize()
    per_step = sorted(events[j].elapsed_time(events[j + 1]) for j in range(KL))
    from airwave_b200 import sharding
    elapsed_ms_max = sharding.max_over_ranks(elapsed_ms) if dist is not None else elapsed_ms
    value = world * n * K * (kb * B / FS) / (elapsed_ms_max * 1e-3)

    # per-kernel pass (same steps, events around every launch) -> roofline of the dominant kernel
    prof_steps = min(K, 200)
    eng.profile_begin(prof_steps)
    for j in range(prof_steps):
        step(W + K + KL + j)
    prof = eng.profile_end()
    peak, peak_src = measured_peaks()
    plan = eng.plan()
    step_bytes = kb * n * algorithmic_bytes(S, B, P, eq_filters)
    step_ms = elapsed_ms_max / K
    kernels_ms = {k: v["ms"] / max(v["launches"], 1) for k, v in prof.items()}
    if len(plan["kernels"]) == 1:
        # one kernel does the whole block (K2+K3+K4): its algorithmic bytes are SURVEY.md 8(d)'s per-stream figure x streams
        dom_name = plan["kernels"][0]
        dom_ms, dom_bytes = kernels_ms[dom_name], kb * n * algorithmic_bytes(S, B, P)
        if eq_filters == 0 and len(banks) == 1:
            # the timed region is K launches of this one kernel and nothing else: its average launch duration is at most the step
            # time (the per-launch events of the profiling pass add a few microseconds of their own)
            dom_ms = min(dom_ms, step_ms)
    else:
        # three kernels per block: the roofline is the slowest one's, with the bytes that kernel has to move (rows = delay lines
        # per stream: speakers that share a filter pair share one)
        rows = banks[0].rows if len(banks) == 1 else S
        split_bytes = {
            0: n * (4 * S * B + 4 * rows * B + 8 * rows * B + 4 * rows * B),   # K2: input block + overlap block in; head-slot spectrum + overlap block out
            1: n * (8 * rows * B * P + 16 * B),                                # K3: every FDL slot read once; two ear accumulators written
            2: n * (16 * B + 8 * B),                                           # K4: accumulators in; two output channels out
        }
        dom_idx = max(range(3), key=lambda i: kernels_ms.get(plan["kernels"][i], 0.0))
        dom_name = plan["kernels"][dom_idx]
        dom_ms = kernels_ms[dom_name]
        dom_bytes = split_bytes[dom_idx]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # the committed ncu capture is of the default call size (profiles/*_traffic.json; "<W>k1" = one block per launch)
    traffic_key = args.workload if kb == e2e_frames // B else (args.workload + "k1" if kb == 1 else None)
    traffic = recorded_traffic(traffic_key, dom_name) if traffic_key else None
    # end-to-end: pinned host buffers, every step's input H2D and output D2H inside the timed region
    F = e2e_frames
    n_bufs = 3
    hin = [aw.PinnedBuffer((n, S, F)) for _ in range(n_bufs)]
    hout = [aw.PinnedBuffer((n, 2, F)) for _ in range(n_bufs)]
    stage = torch.empty((n, S, F), dtype=torch.float32, device=f"cuda:{local}")
    for i, b in enumerate(hin):
        aw._lib.check(aw.lib().aw_synth_fill_device(local, stage.data_ptr(), rank * n, n, S, i * F, F, SEED, None))
        torch.cuda.synchronize()
        b.array[...] = stage.cpu().numpy()
    del stage
    e2e_steps = max(3, min(args.e2e_steps, 10_000))
    for i in range(2):
        eng.submit(hin[i % n_bufs].array.ctypes.data, hout[i % n_bufs].array.ctypes.data, F)
    eng.wait()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.submit(hin[i % n_bufs].array.ctypes.data, hout[i % n_bufs].array.ctypes.data, F)
    eng.wait()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    checksum = float(hout[(e2e_steps - 1) % n_bufs].array[:, :, -1].astype("float64").sum())
    e2e_s_max = sharding.max_over_ranks(e2e_s) if dist is not None else e2e_s
    e2e_value = world * n * e2e_steps * (F / FS) / e2e_s_max
    e2e_checks = [checksum]
    if dist is not None:
        tc = torch.tensor([checksum], dtype=torch.float64, device=f"cuda:{local}")
        allc = [torch.zeros_like(tc) for _ in range(world)]
        dist.all_gather(allc, tc)
        e2e_checks = [float(c.item()) for c in allc]

    # the same metric with ONE block per call (one launch per B-frame block: the real-time contract, a callback brings one block):
    # the latency-oriented figure next to the headline, whose step is the e2e call size
    single = None
    if kb > 1 and args.single_block:
        def sstep(j: int):
            eng.process_device(xp + 4 * (j % R) * B, S * R * B, R * B, yp + (j & 1) * y_alt, 2 * kb * B, kb * B, B)

        Ks = max(20, min(K * kb, 2000))
        for j in range(W):
            sstep(j)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sev = [torch.cuda.Event(enable_timing=True) for _ in range(Ks + 1)]
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for j in range(Ks):
            sstep(j)
        if overlap_eq:
            eng.flush()
        m1.record(stream)
        torch.cuda.synchronize()
        s_ms = m0.elapsed_time(m1)
        s_ms = sharding.max_over_ranks(s_ms) if dist is not None else s_ms
        sev[0].record(stream)
        for j in range(Ks):
            sstep(j)
            sev[j + 1].record(stream)
        torch.cuda.synchronize()
        per_block = sorted(sev[j].elapsed_time(sev[j + 1]) for j in range(Ks))
        s_bytes = n * algorithmic_bytes(S, B, P, eq_filters)
        single = {"blocks_per_call": 1, "calls": Ks, "ms_per_block": s_ms / Ks, "value": world * n * Ks * (B / FS) / (s_ms * 1e-3), "unit": UNIT,
                  "frac": s_bytes / (s_ms / Ks * 1e-3) / 1e9 / peak,
                  "latency_ms": {"p50": per_block[len(per_block) // 2], "p99": percentile(per_block, 0.99), "max": per_block[-1]}}

    shard = None
    if world > 1 or args.shard_check:
        shard = shard_check(aw, torch, dist, rank, world, local, min(n, 512), S, B, (pcm, rate, FS, l_idx, r_idx, B), 4 * B)
    sync_latency = run_sync_latency(aw, local) if (args.workload == "C1" and rank == 0 and world == 1) else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_eq = cpu_eq_definition(args.workload)
        v, cores, sample = cpu_sample(S, B, pcm, rate, l_idx, r_idx, args.cpu_seconds, eq_definition=cpu_eq)   # (first bank of the workload)
        v1, _, sample1 = cpu_sample(S, B, pcm, rate, l_idx, r_idx, min(3.0, args.cpu_seconds), threads=1, eq_definition=cpu_eq)   # the
                                                                                       # reference's own model: one real-time thread
        if cpu_eq is not None:
            sample += "; full chain (convolution + the 10-biquad Double cascade per stream)"
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "one_thread": {"value": v1, "unit": UNIT, "sample": sample1}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "streams_per_gpu": n, "speakers": S, "block": B, "partitions": P,
                       "taps": taps, "fdl_rows": bank.rows, "blocks_per_step": kb, "step": f"{kb} x {B}-frame block(s) for all streams, one call (forward FFT -> FDL multiply-accumulate -> inverse FFT"
                               + (f" -> {eq_filters}-biquad float64 EQ cascade)" if eq_filters else ")"),
                       "l2": f"inputs larger than L2: the FDL working set read every step is {n * 8 * S * B * P / 1e6:.0f} MB (L2 = 126 MB)",
                       "plan": plan, "equalizer": ("overlapped with the next call's convolution (AW_ENGINE_OVERLAP_EQ)" if overlap_eq else
                                                 ("serial" if eq_filters or args.workload == "F3" else None))},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms, "blocks_per_launch": kb,
                         "fdl_rows": bank.rows, "frac_counting_shared_rows":
                             kb * n * (8 * bank.rows * B * P + 4 * S * B + 8 * B) / (dom_ms * 1e-3) / 1e9 / peak if len(plan["kernels"]) == 1 else None,
                         "hbm_minimum_bytes_per_launch": n * call_minimum_bytes(S, B, P, kb),
                         "frac_of_peak_by_hbm_minimum": n * call_minimum_bytes(S, B, P, kb) / (dom_ms * 1e-3) / 1e9 / peak,
                         "note": ("three-kernel path (B < 64, B = 4096): `kernel` is the slowest of K2/K3/K4 and `achieved` the bytes that kernel "
                                  "has to move (delay lines counted as kept: fdl_rows per stream) / its duration; the whole block against "
                                  "SURVEY.md 8(d)'s bytes is step_roofline. " if len(plan["kernels"]) > 1 else "") +
                                 "achieved = SURVEY.md 8(d) bytes per stream per block x streams x blocks of one launch / its duration. A launch "
                                 "that walks k blocks tile-major re-reads a tile's FDL rows from L2, so DRAM `traffic` is below the algorithmic "
                                 "bytes and `frac` can pass 1; hbm_minimum = the history read once + k new slots + input + output. "
                                 "fdl_rows: speakers that share a filter pair (FC and LFE) share one delay line, so the kernel keeps 7 rows "
                                 "per 7.1 stream where the formula counts 8; frac_counting_shared_rows uses 8*rows*B*P for the FDL term"},
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                              "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak, "kernels_ms": kernels_ms},
            "latency_ms": {"p50": per_step[len(per_step) // 2], "p99": per_step[min(len(per_step) - 1, int(0.99 * len(per_step)))],
                           "max": per_step[-1], "per": f"call of {kb} block(s) = {kb * B} frames for all streams",
                           "p50_per_block": per_step[len(per_step) // 2] / kb, "p99_per_block": percentile(per_step, 0.99) / kb},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * S * F * 4, "d2h_bytes_per_step": n * 2 * F * 4,
                    "frames_per_step": F, "steps": e2e_steps, "timing": "host wall clock around aw_engine_submit..aw_engine_wait, "
                    "synchronised on both sides, max over ranks; copies overlap kernels across steps (2 staging sets)",
                    "checksum": checksum, "per_rank_checksum": e2e_checks},
            "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
            "host_affinity": affinity,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if single is not None:
            line["single_block_calls"] = single
        if shard is not None:
            line["sharding"] = shard
        if sync_latency is not None:
            line["sync_latency"] = sync_latency
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--blocks-per-call", type=int, default=0,
                    help="blocks per device-resident call of the timed region (a step); 0 = the e2e call size, --e2e-frames / block")
    ap.add_argument("--serial-eq", action="store_true", help="C4/F3: run the equalizer behind its call's convolution on the same stream")
    ap.add_argument("--no-single-block", dest="single_block", action="store_false",
                    help="skip the secondary measurement with one block per call")
    ap.add_argument("--shard-check", action="store_true", help="run the sharding bit-identity check even on one GPU")
    ap.add_argument("--offline-blocks", default="64,128,256,512,1024,2048,4096", help="C5-offline: block sizes of the sweep")
    ap.add_argument("--offline-seconds", type=float, default=60.0, help="C5-offline: seconds of audio rendered per stream")
    ap.add_argument("--e2e-frames", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "C5-offline":
        return run_offline(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
