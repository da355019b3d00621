"""GPU stress (not part of the test-suite): full-size twin check over long runs, for calls of 1 block, 4 blocks and 4096 frames —
every stream bit-identical to its twin, every call size bit-identical to the others, one stream against float64 convolution.
usage: python tools/stress_round2.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import airwave_b200 as aw
import oracle

FS = 48000.0
SEED = 0x41495257
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run(n, B, blocks, unique=3, S=8):
    lay = aw.InputLayout.surround71()
    wav = aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", "RoomSH1.0.wav"))
    bank = aw.HRIRBank.from_wav(wav, FS, lay, B)
    frames = blocks * B
    xu = oracle.synth_block(SEED, [11 + 5 * i for i in range(unique)], S, 0, frames)
    reps = -(-n // unique)
    base = None
    for per_call in sorted({B, min(4 * B, 4096), 4096}):
        eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=per_call, max_partitions=bank.partitions)
        eng.set_bank(bank)
        outs = []
        for a in range(0, frames, per_call):
            outs.append(eng.process(np.ascontiguousarray(np.tile(xu[:, :, a:a + per_call], (reps, 1, 1))[:n])))
        y = np.concatenate(outs, axis=2)
        bad = sum(1 for i in range(unique, n) if not np.array_equal(y[i], y[i % unique]))
        same = True if base is None else bool(np.array_equal(y[:unique], base))
        if base is None:
            base = y[:unique].copy()
            h = oracle.hrir_matrix(oracle.load_wav(os.path.join(GOLDEN, "hrtf", "RoomSH1.0.wav")), FS, oracle.InputLayout.surround71)
            err = float(np.abs(y[0] - oracle.direct_conv_f64(xu[0], h)).max())
        print(f"n={n} B={B} blocks={blocks} frames/call={per_call} ({per_call // B} blocks/launch) {eng.plan()['kernels']}: "
              f"{bad} streams differ from their twin; identical to 1-block calls: {same}; max-abs vs float64 {err:.2e}", flush=True)
        eng.close()
        assert bad == 0 and same and err <= 1e-5


run(4096, 256, 256)
run(2048, 64, 512)
run(2048, 128, 384)
run(2048, 512, 128)
run(2048, 1024, 48)
run(2048, 2048, 24)
run(8192, 256, 64)
run(3001, 256, 96)
run(777, 64, 256)
print("stress ok")
