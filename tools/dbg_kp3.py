"""Which (frame, partition) product explains the error of block 0?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import airwave_b200 as aw
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, n, kb, blocks = 64, 23, 2, 8
lay = aw.InputLayout.surround71()
bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(os.path.join(GOLDEN, "hrtf", "StageSH1.0.wav")), 48000.0, lay, B)
P = bank.partitions
h = oracle.hrir_matrix(oracle.load_wav(os.path.join(GOLDEN, "hrtf", "StageSH1.0.wav")), 48000.0, oracle.InputLayout.surround71)  # [S][2][taps]
hp = np.zeros((8, 2, P * B)); hp[:, :, :h.shape[2]] = h
x = oracle.synth_block(0x41495257, [101 + 7 * i for i in range(n)], 8, 0, blocks * B).astype(np.float64)
for trial in range(6):
    eng = aw.BinauralEngine(n, 8, B, 48000.0, max_frames_per_call=kb * B, max_partitions=P)
    eng.set_bank(bank)
    y = np.concatenate([eng.process(np.ascontiguousarray(x[:, :, a:a + kb * B].astype(np.float32))) for a in range(0, blocks * B, kb * B)], axis=2)
    eng.close()
    for i in (0, 11):
        ref = oracle.direct_conv_f64(x[i].astype(np.float32), h)
        for blk in range(blocks):
            e = (y[i] - ref)[:, blk * B:(blk + 1) * B]
            if np.abs(e).max() < 1e-5:
                continue
            best = []
            for fb in range(0, min(blocks, blk + 3)):          # frame [x_{fb-1} | x_fb]
                prev = x[i][:, (fb - 1) * B:fb * B] if fb > 0 else np.zeros((8, B))
                frame = np.concatenate([prev, x[i][:, fb * B:(fb + 1) * B]], axis=1)
                F = np.fft.rfft(frame, axis=1)
                for p in range(P):
                    H = np.fft.rfft(np.concatenate([hp[:, :, p * B:(p + 1) * B], np.zeros((8, 2, B))], axis=2), axis=2)
                    c = np.fft.irfft((F[:, None, :] * H).sum(axis=0), axis=1)[:, B:]
                    for sign in (1, -1):
                        r = np.abs(e - sign * c).max()
                        best.append((r, fb, p, sign))
            best.sort()
            print(f"trial {trial} stream {i} block {blk}: |e| {np.abs(e).max():.3e}; best explanations (residual, frame, partition, sign): {[(float('%.2e' % b[0]),) + b[1:] for b in best[:3]]}")
