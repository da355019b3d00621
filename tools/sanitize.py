"""GPU-side helper: a few small renders through every execution path, meant to be run under compute-sanitizer
(memcheck / synccheck / initcheck).  usage: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import airwave_b200 as aw

FS = 48000.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wav = aw.WAVLoader.load(os.path.join(root, "tests", "golden", "hrtf", "RoomSH1.0.wav"))
eqtxt = open(os.path.join(root, "tests", "golden", "eq", "CCA CRA ParametricEq.txt"), "rb").read()
definition = aw.EqualizerAPOParser.parse(eqtxt, "f")
rng = np.random.default_rng(0)
for block, n, env in [(64, 9, {}), (128, 9, {}), (256, 21, {}), (512, 9, {}), (1024, 5, {}), (2048, 3, {}), (4096, 2, {}),
                      (256, 9, {"AW_PERSISTENT": "0"}), (256, 9, {"AW_FUSED_TILE": "0"}), (256, 9, {"AW_EQ_FUSION": "1"}),
                      (256, 13, {"AW_PERSISTENT_CTAS": "2"}), (256, 700, {}), (512, 650, {}), (64, 640, {}), (1024, 301, {}),
                      # round 2: bulk-copy fallback, block-major walk, one launch per block, staged copies instead of zero-copy,
                      # the reference's ring modulus
                      (256, 21, {"AW_KP_TENSOR_TMA": "0"}), (64, 37, {"AW_KP_TENSOR_TMA": "0"}), (128, 300, {"AW_KP_ORDER": "0"}),
                      (256, 9, {"AW_KP_MULTIBLOCK": "0"}), (64, 9, {"AW_ZERO_COPY": "0"}), (512, 9, {"AW_KP_RING_EXTRA": "0"}),
                      # the stand-alone transforms: small blocks, CTAs that loop over frames (more frames than resident CTAs), one
                      # delay line per speaker, the fused kernel with shared rows at another block size
                      (32, 9, {}), (2048, 70, {"AW_FUSED_TILE": "0", "AW_SA_WAVES": "1"}), (4096, 3, {"AW_KP_MERGE_ROWS": "0"}),
                      (512, 9, {"AW_PERSISTENT": "0"})]:
    # AW_SANITIZE_ONLY=transforms: only the plans that run K2 / K4 / KF (what a change to those kernels needs re-checked)
    if os.environ.get("AW_SANITIZE_ONLY") == "transforms" and not (block in (32, 4096) or "AW_FUSED_TILE" in env or "AW_PERSISTENT" in env):
        continue
    os.environ.update(env)
    bank = aw.HRIRBank.from_wav(wav, FS, aw.InputLayout.surround71(), block)
    eng = aw.BinauralEngine(n, 8, block, FS, max_frames_per_call=min(4096, 2 * block))
    for k in env:
        os.environ.pop(k)
    eng.set_bank(bank)
    eng.eq_prepare(definition)
    for call in range(3):
        x = rng.uniform(-0.25, 0.25, (n, 8, min(4096, 2 * block))).astype(np.float32)
        y = eng.process(x)
    y = eng.process(x[:, :, :37])          # ragged call: adapter path
    assert np.isfinite(y).all()
    print("ok", block, n, env, eng.plan()["kernels"], flush=True)
    eng.close()
