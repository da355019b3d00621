"""Debug helper (GPU): where do twin streams first differ?  usage: python tools/dbg_twins.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import airwave_b200 as aw

FS = 48000.0
SEED = 0x41495257

def run(n, B, taps, blocks, per_call, unique=3, S=8, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    rng = np.random.default_rng(1)
    t = np.arange(taps)
    pcm = (0.05 * rng.standard_normal((14, taps)) * np.exp(-t / (0.25 * FS))).astype(np.float32)
    pcm[:, min(190, taps - 1)] += 0.5
    lay = aw.InputLayout.surround71()
    m = aw.HRIRChannelMap.hesuvi14Channel(lay.channels)
    l = [m.getIndices(s)[0] for s in lay.channels]; r = [m.getIndices(s)[1] for s in lay.channels]
    bank = aw.HRIRBank(pcm, FS, FS, l, r, B)
    eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=per_call, max_partitions=bank.partitions)
    eng.set_bank(bank)
    frames = blocks * B
    xu = np.random.default_rng(SEED).uniform(-0.25, 0.25, (unique, S, frames)).astype(np.float32)
    reps = -(-n // unique)
    outs = []
    for a in range(0, frames, per_call):
        outs.append(eng.process(np.ascontiguousarray(np.tile(xu[:, :, a:a + per_call], (reps, 1, 1))[:n])))
    y = np.concatenate(outs, axis=2)
    bad = []
    for i in range(unique, n):
        d = np.abs(y[i] - y[i % unique])
        if d.max() > 0:
            idx = np.argwhere(d > 0)
            bad.append((i, int(idx[:, 1].min()) // B, int(idx[:, 1].max()) // B, float(d.max()), int((d > 0).sum())))
    print(f"n={n} B={B} P={bank.partitions} blocks={blocks} per_call={per_call} env={env} plan={eng.plan()['kernels']}: {len(bad)} bad streams", flush=True)
    for b in bad[:12]:
        print("   stream %d: first bad block %d, last %d, max diff %.3e, %d samples" % b)
    for k in (env or {}):
        os.environ.pop(k, None)
    eng.close()

run(1024, 512, 65536, 132, 4096)
run(1024, 512, 65536, 40, 4096)
run(1024, 512, 65536, 40, 512)
run(296, 512, 65536, 40, 4096)
run(1024, 512, 4320, 40, 4096)
run(1024, 512, 65536, 40, 4096, env={"AW_PERSISTENT_TILE": "2"})
run(1024, 256, 32768, 40, 2048)
