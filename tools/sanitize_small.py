"""racecheck subset: one small render per kernel family (see tools/sanitize.py for the full memcheck sweep)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import airwave_b200 as aw
FS = 48000.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wav = aw.WAVLoader.load(os.path.join(root, "tests", "golden", "hrtf", "RoomSH1.0.wav"))
definition = aw.EqualizerAPOParser.parse(open(os.path.join(root, "tests", "golden", "eq", "CCA CRA ParametricEq.txt"), "rb").read(), "f")
rng = np.random.default_rng(0)
for block, n, env in [(64, 20, {"AW_PERSISTENT_CTAS": "2"}), (256, 24, {"AW_PERSISTENT_CTAS": "2", "AW_PERSISTENT_TILE": "4"}), (512, 12, {"AW_PERSISTENT_CTAS": "2", "AW_PERSISTENT_TILE": "4"}), (1024, 5, {}),
                      (4096, 3, {}), (32, 9, {}), (256, 9, {"AW_PERSISTENT": "0"}), (2048, 40, {"AW_FUSED_TILE": "0", "AW_SA_WAVES": "1"})]:
    if os.environ.get("AW_SANITIZE_ONLY") == "transforms" and not (block in (32, 4096) or "AW_FUSED_TILE" in env or "AW_PERSISTENT" in env):
        continue
    os.environ.update(env)
    bank = aw.HRIRBank.from_wav(wav, FS, aw.InputLayout.surround71(), block)
    eng = aw.BinauralEngine(n, 8, block, FS, max_frames_per_call=2 * block)
    for k in env:
        os.environ.pop(k)
    eng.set_bank(bank)
    eng.eq_install_state(definition)
    for call in range(2):
        y = eng.process(rng.uniform(-0.25, 0.25, (n, 8, 2 * block)).astype(np.float32))
    assert np.isfinite(y).all()
    print("ok", block, n, eng.plan()["kernels"], flush=True)
    eng.close()
