"""GPU-side debugging aid: per-block error of the persistent kernel against float64 direct convolution for a few plans."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def run(B, n, per_call_blocks, blocks, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    import airwave_b200 as aw
    lay = aw.InputLayout.surround71()
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", "StageSH1.0.wav")), 48000.0, lay, B)
    eng = aw.BinauralEngine(n, 8, B, 48000.0, max_frames_per_call=per_call_blocks * B, max_partitions=bank.partitions)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    eng.set_bank(bank)
    x = oracle.synth_block(0x41495257, [101 + 7 * i for i in range(n)], 8, 0, blocks * B)
    pc = per_call_blocks * B
    y = np.concatenate([eng.process(np.ascontiguousarray(x[:, :, a:a + pc])) for a in range(0, blocks * B, pc)], axis=2)
    h = oracle.hrir_matrix(oracle.load_wav(os.path.join(GOLDEN, "hrtf", "StageSH1.0.wav")), 48000.0, oracle.InputLayout.surround71)
    worst = []
    for i in sorted(set([0, 1, n // 2, n - 1])):
        ref = oracle.direct_conv_f64(x[i], h)
        err = np.abs(y[i] - ref).reshape(2, blocks, B).max(axis=(0, 2))
        bad = [int(b) for b in np.nonzero(err > 1e-5)[0]]
        e0 = (y[i] - ref)[0, :B] if bad else None
        alt = None
        if bad:
            bb = bad[0]
            e = (y[i] - ref)[0, bb * B:(bb + 1) * B]
            sign = np.where(np.arange(B) % 2 == 0, 1.0, -1.0)
            alt = (float(np.abs(e).max()), float(np.abs(e - e.mean()).max()), float(np.abs(e - sign * (e * sign).mean()).max()))
        worst.append((i, float(err.max()), bad[:12], alt))
    print(f"B={B} n={n} blocks/call={per_call_blocks} env={env} plan={eng.plan()['kernels']}")
    for w in worst:
        print("   stream %d max %.3e bad blocks %s (max, after removing DC, after removing Nyquist) %s" % w)
    eng.close()

for B in (64, 128):
    P = -(-4320 // B)
    for env in ({}, {"AW_KP_MULTIBLOCK": "0"}, {"AW_PERSISTENT_TILE": "2"}, {"AW_PERSISTENT_CTAS": "1"}, {"AW_ZERO_COPY": "0"}):
        run(B, 23, 2, 12, env)
    run(B, 23, 1, 12, {})
    run(B, 4, 2, 12, {})
    run(B, 300, 2, 8, {})
run(512, 3, 4, 12, {})
run(1024, 2, 4, 8, {})
