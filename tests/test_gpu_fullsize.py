"""GPU parity at BASELINE.json's full sizes, and agreement between the execution paths, through the C ABI.

At 4,096 / 8,192 / 1,024 / 2,048 streams an element-wise CPU comparison of every stream is too slow for the oracle, so
the tests use a size-independent property the domain offers — streams are independent and the engine is deterministic —
on top of the oracle: the batch carries U distinct input signals repeated over all streams (so every tile of every
persistent CTA sees all of them at every position inside a tile), the U distinct outputs are checked against the float64
direct-convolution oracle (max-abs 1e-5, SNR >= 100 dB: BASELINE.json north_star), and every other stream must be
BIT-IDENTICAL to its twin.  A stream whose rows were mixed up with a neighbour's, a tile that was skipped or processed
twice, a ring slot that was read one block early — all of these break either the oracle check or the twin check.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, snr_db

pytestmark = pytest.mark.gpu

MAX_ABS = 1e-5
SNR_DB = 100.0
SEED = 0x41495257
FS = 48000.0


@pytest.fixture(scope="module")
def aw():
    import airwave_b200
    assert airwave_b200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return airwave_b200


def _maps(aw, layout):
    m = aw.HRIRChannelMap.hesuvi14Channel(layout.channels)
    return [m.getIndices(s)[0] for s in layout.channels], [m.getIndices(s)[1] for s in layout.channels]


def _render_twins(aw, bank, n, S, B, blocks, unique, per_call, pcm_lr, eq=None, env=None):
    """Renders n streams carrying `unique` distinct signals (stream i carries signal i % unique); returns (x_unique, y)."""
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        eng = aw.BinauralEngine(n, S, B, FS, max_frames_per_call=per_call, max_partitions=bank.partitions)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    eng.set_bank(bank)
    if eq is not None:
        eng.eq_prepare(eq)
    frames = blocks * B
    xu = oracle.synth_block(SEED, [101 + 7 * i for i in range(unique)], S, 0, frames)
    reps = -(-n // unique)
    outs = []
    for a in range(0, frames, per_call):
        chunk = np.tile(xu[:, :, a:a + per_call], (reps, 1, 1))[:n]
        outs.append(eng.process(np.ascontiguousarray(chunk)))
    y = np.concatenate(outs, axis=2)
    plan = eng.plan()
    eng.close()
    return xu, y, plan


def _check_twins_and_oracle(xu, y, h, unique, eq_definition=None):
    n = y.shape[0]
    for i in range(unique, n):                       # every stream is bit-identical to its twin among the first `unique`
        assert np.array_equal(y[i], y[i % unique]), f"stream {i} differs from its twin {i % unique}"
    for i in range(unique):
        ref = oracle.direct_conv_f64(xu[i], h)
        if eq_definition is not None:
            # float64 convolution narrowed to Float (RealtimeAudioProcessor output), then the reference's Double biquad cascade
            o = oracle.ParametricEqualizerProcessor(FS)
            o.setTarget(eq_definition)
            l, r = [], []
            for a in range(0, ref.shape[1], 4096):
                ll, rr = o.process(ref[0, a:a + 4096].astype(np.float32), ref[1, a:a + 4096].astype(np.float32))
                l.append(ll); r.append(rr)
            ref = np.stack([np.concatenate(l), np.concatenate(r)]).astype(np.float64)
        assert np.abs(y[i] - ref).max() <= MAX_ABS, i
        assert snr_db(ref, y[i]) >= SNR_DB, i


def test_c2_full_size_4096_streams(aw, hrtf_path):
    """BASELINE configs[1]: 7.1 -> binaural, RoomSH1.0, B = 256, 4,096 concurrent streams on one GPU."""
    lay = aw.InputLayout.surround71()
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, lay, 256)
    xu, y, plan = _render_twins(aw, bank, 4096, 8, 256, blocks=24, unique=12, per_call=1024, pcm_lr=None)
    assert plan["kernels"] == ["k_persistent<8,4>"]
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), FS, oracle.InputLayout.surround71)
    _check_twins_and_oracle(xu, y, h, 12)


def test_c4_full_chain_8192_streams(aw, hrtf_path, eq_fixture_bytes):
    """BASELINE configs[3]: 7.1 HRIR convolution + 10-band parametric EQ + HRIR 44.1 -> 48 kHz resample, 8,192 streams.
    The bundled presets are 48 kHz files, so the 44.1 kHz source is StageSH1.0 relabelled."""
    lay = aw.InputLayout.surround71()
    wav_o = oracle.load_wav(hrtf_path("StageSH1.0"))
    l, r = _maps(aw, lay)
    bank = aw.HRIRBank(wav_o.audioData, 44100.0, FS, l, r, 256)
    assert (bank.taps, bank.partitions) == (4702, 19)
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    assert len(definition["filters"]) == 10
    # installing the state directly (no 20 ms crossfade from unity) keeps the reference chain simple: conv -> cascade
    old = None
    n, unique = 8192, 10
    eng = aw.BinauralEngine(n, 8, 256, FS, max_frames_per_call=1024, max_partitions=19)
    eng.set_bank(bank)
    eng.eq_install_state(definition)
    frames = 24 * 256
    xu = oracle.synth_block(SEED, [3 + 5 * i for i in range(unique)], 8, 0, frames)
    reps = -(-n // unique)
    y = np.concatenate([eng.process(np.ascontiguousarray(np.tile(xu[:, :, a:a + 1024], (reps, 1, 1))[:n]))
                        for a in range(0, frames, 1024)], axis=2)
    assert eng.plan()["kernels"] == ["k_persistent<8,4>"]
    eng.close()
    wav_o.sampleRate = 44100.0
    h = oracle.hrir_matrix(wav_o, FS, oracle.InputLayout.surround71)
    assert h.shape[2] == 4702
    for i in range(unique, n):
        assert np.array_equal(y[i], y[i % unique]), i
    # (1) against the reference's own arithmetic — the float32 restatement of RealtimeAudioProcessor/ConvolutionEngine followed by
    #     the Double biquad cascade (AudioEffectGraph.swift:195-210): BASELINE.json's bound, unscaled
    # (2) against float64 direct convolution followed by the same cascade: the same bound, unscaled (measured on B200:
    #     max-abs 1.8e-7, SNR 133 dB — profiles/r02_c4_chain_error.json)
    rows = []
    for i in range(unique):
        rap = oracle.RealtimeAudioProcessor(oracle.activate_preset(wav_o, FS, oracle.InputLayout.surround71, 256), 256, 256, literalStereo=False)
        conv32 = np.concatenate([np.stack(rap.process_channels([xu[i, sp, b * 256:(b + 1) * 256] for sp in range(8)]))
                                 for b in range(frames // 256)], axis=1)
        st32 = oracle.ParametricEqualizerState(definition, FS)
        ref32 = np.stack(st32.process(conv32[0], conv32[1])).astype(np.float64)
        e32, s32 = float(np.abs(y[i] - ref32).max()), float(snr_db(ref32, y[i]))
        assert e32 <= MAX_ABS and s32 >= SNR_DB, (i, e32, s32)
        conv = oracle.direct_conv_f64(xu[i], h)
        st = oracle.ParametricEqualizerState(definition, FS)
        el, er = st.process(conv[0].astype(np.float32), conv[1].astype(np.float32))
        ref = np.stack([el, er]).astype(np.float64)
        scale = max(1.0, float(np.abs(ref).max() / max(np.abs(conv).max(), 1e-30)))
        e64, s64 = float(np.abs(y[i] - ref).max()), float(snr_db(ref, y[i]))
        rows.append(dict(stream=i, max_abs_vs_float32_reference_chain=e32, snr_db_vs_float32_reference_chain=s32,
                         max_abs_vs_float64_chain=e64, snr_db_vs_float64_chain=s64, eq_peak_gain=scale))
        assert e64 <= MAX_ABS and s64 >= SNR_DB, (i, e64, s64)
    evidence = os.environ.get("AW_EVIDENCE_DIR")
    if evidence and os.path.isdir(evidence):
        import json
        with open(os.path.join(evidence, "c4_chain_error.json"), "w") as f:
            json.dump(dict(workload="C4 full chain, 8192 streams, 24 blocks of 256 frames", bound_max_abs=MAX_ABS, bound_snr_db=SNR_DB, streams=rows), f, indent=1)


def test_c3_full_size_1024_streams_long_brir(aw):
    """BASELINE configs[2]: synthetic 65,536-tap BRIR, P = 128 at B = 512, 1,024 streams; more than P blocks so every
    stream's ring wraps (Q4)."""
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(1)
    taps, B, S = 65536, 512, 8
    t = np.arange(taps)
    pcm = (0.05 * rng.standard_normal((14, taps)) * np.exp(-t / (0.25 * FS))).astype(np.float32)
    pcm[:, 190] += 0.5
    lay = aw.InputLayout.surround71()
    l, r = _maps(aw, lay)
    bank = aw.HRIRBank(pcm, FS, FS, l, r, B)
    assert bank.partitions == 128
    unique, blocks = 3, 132
    xu, y, plan = _render_twins(aw, bank, 1024, S, B, blocks=blocks, unique=unique, per_call=4096, pcm_lr=None)
    assert plan["kernels"][0].startswith("k_persistent<9,")
    for i in range(unique, 1024):
        assert np.array_equal(y[i], y[i % unique]), i
    for i in range(unique):
        ref = np.zeros((2, blocks * B))
        for s in range(S):
            ref[0] += fftconvolve(xu[i, s].astype(np.float64), pcm[l[s]].astype(np.float64))[: blocks * B]
            ref[1] += fftconvolve(xu[i, s].astype(np.float64), pcm[r[s]].astype(np.float64))[: blocks * B]
        assert np.abs(y[i] - ref).max() <= MAX_ABS, i
        assert snr_db(ref, y[i]) >= SNR_DB, i


@pytest.mark.parametrize("block", [64, 128, 256, 512, 1024, 2048, 4096])
def test_c5_block_size_sweep_2048_streams(aw, hrtf_path, block):
    """BASELINE configs[4] per GPU: 2,048 streams, block sizes 64 ... 4096 (P = 68 ... 2)."""
    lay = aw.InputLayout.surround71()
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, lay, block)
    P = -(-4320 // block)
    assert bank.partitions == P
    blocks = max(3, min(P + 3, 4 * 4096 // block))
    per_call = min(4096, 4 * block)
    xu, y, plan = _render_twins(aw, bank, 2048, 8, block, blocks=blocks, unique=6, per_call=per_call, pcm_lr=None)
    if 64 <= block <= 2048:
        assert plan["kernels"][0].startswith("k_persistent<"), plan
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), FS, oracle.InputLayout.surround71)
    _check_twins_and_oracle(xu, y, h, 6)


PATHS = [  # (name, env, block sizes it exists for)
    ("persistent T=4", {"AW_PERSISTENT_TILE": "4"}, (64, 128, 256, 512)),
    ("persistent T=2", {"AW_PERSISTENT_TILE": "2"}, (64, 128, 256, 512, 1024, 2048)),
    ("persistent T=2, 3 CTAs", {"AW_PERSISTENT_CTAS": "3", "AW_PERSISTENT_TILE": "2"}, (64, 128, 256, 512, 1024, 2048)),
    ("persistent T=4, 2 CTAs", {"AW_PERSISTENT_CTAS": "2", "AW_PERSISTENT_TILE": "4"}, (64, 128, 256, 512)),
    ("persistent, one launch per block", {"AW_KP_MULTIBLOCK": "0"}, (64, 128, 256, 512, 1024, 2048)),
    ("persistent, block-major", {"AW_KP_ORDER": "0", "AW_PERSISTENT_CTAS": "2"}, (64, 128, 256, 512, 1024, 2048)),
    ("persistent, tile-major, 2 CTAs", {"AW_KP_ORDER": "1", "AW_PERSISTENT_CTAS": "2", "AW_KP_KEEP": "100"}, (64, 128, 256, 512, 1024, 2048)),
    ("fused", {"AW_PERSISTENT": "0"}, (64, 128, 256, 512)),
    ("split", {"AW_FUSED_TILE": "0"}, (64, 128, 256, 512, 1024, 2048)),
]


@pytest.mark.parametrize("block", [64, 128, 256, 512, 1024, 2048])
def test_execution_paths_agree(aw, hrtf_path, block):
    """KP (both tile sizes, few CTAs so that every CTA walks many tiles), KF and the split kernels K2/K3/K4 compute the same
    convolution; they differ only in summation order, so they agree to float32 rounding and each meets the oracle bound.
    Odd stream counts exercise the partial last tile."""
    lay = aw.InputLayout.surround71()
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("StageSH1.0")), FS, lay, block)
    P = bank.partitions
    blocks = max(3, min(P + 2, 40))
    n = 23
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("StageSH1.0")), FS, oracle.InputLayout.surround71)
    results = {}
    for name, env, sizes in PATHS:
        if block not in sizes:
            continue
        xu, y, plan = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=n, per_call=min(4096, 2 * block), pcm_lr=None, env=env)
        results[name] = (y, plan["kernels"])
        for i in (0, 10, 22):
            ref = oracle.direct_conv_f64(xu[i], h)
            assert np.abs(y[i] - ref).max() <= MAX_ABS, (name, i)
            assert snr_db(ref, y[i]) >= SNR_DB, (name, i)
    kernels = {name: k for name, (_, k) in results.items()}
    assert kernels["split"][1].startswith("k_fdl_cmac"), kernels
    if "fused" in kernels:
        assert kernels["fused"][0].startswith("k_fused<"), kernels
    base = results["persistent T=2"][0]
    if "persistent T=4" in results:   # the tile size must not change a single bit (multi-GPU sharding relies on it)
        assert np.array_equal(results["persistent T=4"][0], base)
    for few in ("persistent T=2, 3 CTAs", "persistent T=4, 2 CTAs"):   # every CTA walks several tiles
        if few in results:
            assert np.array_equal(results[few][0], base), few
    # a call of k blocks is one launch that walks (tile, block) items: neither one launch per block nor the walk order
    # (block-major / tile-major) may change a bit
    for variant in ("persistent, one launch per block", "persistent, block-major", "persistent, tile-major, 2 CTAs"):
        assert np.array_equal(results[variant][0], base), variant
    for name, (y, _) in results.items():
        assert np.abs(y - base).max() <= 4e-6, name


@pytest.mark.parametrize("block,n", [(64, 700), (256, 1201), (512, 37), (1024, 301)])
def test_blocks_per_call_do_not_change_a_bit(aw, hrtf_path, block, n):
    """RealtimeAudioProcessor.process cuts a callback into B-frame blocks (RealtimeAudioProcessor.swift:88-116): how a caller
    cuts its audio into calls must not change the samples.  Calls of 1, 3 and 16 blocks (one KP launch each, the ring slots of a
    (stream, speaker) written and read across block boundaries inside the launch) against one block per call, with enough
    streams that every CTA walks several tiles and the last round is cut into partial tiles."""
    lay = aw.InputLayout.surround71()
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, lay, block)
    blocks = 48 if block <= 256 else 12
    unique = 5
    base = None
    for k, env in [(1, None), (3, None), (16, None), (4, {"AW_KP_ORDER": "0"}), (4, {"AW_PERSISTENT_TILE": "2", "AW_KP_KEEP": "0"})]:
        if k * block > 4096:
            k = 4096 // block
        xu, y, plan = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=unique, per_call=k * block, pcm_lr=None, env=env)
        assert plan["kernels"][0].startswith("k_persistent<"), plan
        for i in range(unique, n):
            assert np.array_equal(y[i], y[i % unique]), (k, i)
        if base is None:
            base = y[:unique].copy()
            h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), FS, oracle.InputLayout.surround71)
            ref = oracle.direct_conv_f64(xu[0], h)
            assert np.abs(base[0] - ref).max() <= MAX_ABS and snr_db(ref, base[0]) >= SNR_DB
        else:
            assert np.array_equal(y[:unique], base), (k, env)


def test_small_and_ragged_stream_counts_and_stereo_on_the_persistent_path(aw, hrtf_path):
    """n = 1, 2, 5 (partial tiles), S = 2 (bank narrower than a stage row group) and P = 1 (heads only)."""
    wav_o = oracle.load_wav(hrtf_path("NeutralSH1.0"))
    for layout, S, n, block in [("stereo", 2, 1, 512), ("stereo", 2, 5, 64), ("surround51", 6, 2, 128), ("surround71", 8, 5, 1024)]:
        lay = getattr(aw.InputLayout, layout)()
        bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("NeutralSH1.0")), FS, lay, block)
        blocks = bank.partitions + 2
        xu, y, plan = _render_twins(aw, bank, n, S, block, blocks=blocks, unique=n, per_call=block, pcm_lr=None)
        assert plan["kernels"][0].startswith("k_persistent<"), plan
        h = oracle.hrir_matrix(wav_o, FS, getattr(oracle.InputLayout, layout))
        for i in range(n):
            ref = oracle.direct_conv_f64(xu[i], h)
            assert np.abs(y[i] - ref).max() <= MAX_ABS and snr_db(ref, y[i]) >= SNR_DB, (layout, n, block, i)
    # P = 1 (heads only) and P = 2, 5 (stages that mix history rows with the head row at B <= 128): short responses
    lay = aw.InputLayout.stereo()
    l, r = _maps(aw, lay)
    for taps, block in [(200, 256), (50, 64), (100, 64), (300, 64), (120, 128), (500, 128)]:
        short = wav_o.audioData[:, :taps].copy()
        bank = aw.HRIRBank(short, FS, FS, l, r, block)
        assert bank.partitions == -(-taps // block)
        xu, y, plan = _render_twins(aw, bank, 9, 2, block, blocks=bank.partitions + 4, unique=9, per_call=block, pcm_lr=None)
        assert plan["kernels"][0].startswith("k_persistent<"), plan
        hs = np.stack([np.stack([short[l[s]], short[r[s]]]) for s in range(2)])
        for i in range(9):
            ref = oracle.direct_conv_f64(xu[i], hs)
            assert np.abs(y[i] - ref).max() <= MAX_ABS, (taps, block, i)


def test_fused_equalizer_epilogue_is_bit_identical_to_the_separate_pass(aw, hrtf_path, eq_fixture_bytes):
    """AW_EQ_FUSION=1 (the cascade rides in the block kernel's epilogue) must not change a single bit: same operations in the
    same order per biquad (ParametricEqualizerProcessor.swift:65-90), steady state after a finished crossfade included."""
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "f.txt")
    outs = {}
    for block, n in [(256, 37), (512, 9)]:
        lay = aw.InputLayout.surround71()
        bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, lay, block)
        x = oracle.synth_block(SEED, range(n), 8, 0, 16 * block)
        for fusion in ("0", "1"):
            os.environ["AW_EQ_FUSION"] = fusion
            try:
                eng = aw.BinauralEngine(n, 8, block, FS, max_frames_per_call=4 * block)
            finally:
                os.environ.pop("AW_EQ_FUSION", None)
            eng.set_bank(bank)
            eng.eq_prepare(definition)            # crossfade from unity over the first 960 frames, then steady state
            outs[fusion] = np.concatenate([eng.process(x[:, :, a:a + 4 * block]) for a in range(0, 16 * block, 4 * block)], axis=2)
            eng.close()
        assert np.array_equal(outs["0"], outs["1"]), block


def test_many_profile_ranges_render_concurrently_and_match_single_bank_engines(aw, hrtf_path):
    """per-device profile semantics at batch scale — 24 stream ranges bound alternately to three banks (one grid
    per range, spread over side streams).  Every stream must be bit-identical to the same stream rendered by an engine that
    only knows its bank."""
    lay = aw.InputLayout.surround71()
    names = ["RoomSH1.0", "StageSH1.0", "NeutralSH1.0"]
    banks = [aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path(nm)), FS, lay, 256) for nm in names]
    n, ranges = 1000, 24
    bounds = [round(i * n / ranges) for i in range(ranges + 1)]
    eng = aw.BinauralEngine(n, 8, 256, FS, max_frames_per_call=1024, max_partitions=17)
    for r in range(ranges):
        eng.set_bank(banks[r % 3], bounds[r], bounds[r + 1] - bounds[r])
    x = oracle.synth_block(SEED, range(40), 8, 0, 20 * 256)
    x = np.ascontiguousarray(np.tile(x, (25, 1, 1)))          # stream i carries signal i % 40
    y = np.concatenate([eng.process(x[:, :, a:a + 1024]) for a in range(0, 20 * 256, 1024)], axis=2)
    eng.close()
    for b in range(3):
        ref = aw.BinauralEngine(40, 8, 256, FS, max_frames_per_call=1024)
        ref.set_bank(banks[b])
        yr = np.concatenate([ref.process(x[:40, :, a:a + 1024]) for a in range(0, 20 * 256, 1024)], axis=2)
        ref.close()
        for r in range(b, ranges, 3):
            for i in range(bounds[r], bounds[r + 1]):
                assert np.array_equal(y[i], yr[i % 40]), (r, i)


def test_bank_with_fewer_speakers_than_the_engine_has_channels(aw, hrtf_path):
    """A 5.1 bank on an engine laid out for 8 input channels (state arrays strided by the engine's channel count, renderers =
    the bank's speakers, RealtimeAudioProcessor.swift:145-147): channels 6 and 7 are ignored and the result is bit-identical
    to a 6-channel engine fed the first 6 channels."""
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, aw.InputLayout.surround51(), 256)
    n = 11
    x = oracle.synth_block(SEED, range(n), 8, 0, 24 * 256)
    wide = aw.BinauralEngine(n, 8, 256, FS, max_frames_per_call=1024, max_partitions=17)
    wide.set_bank(bank)
    narrow = aw.BinauralEngine(n, 6, 256, FS, max_frames_per_call=1024, max_partitions=17)
    narrow.set_bank(bank)
    yw = np.concatenate([wide.process(x[:, :, a:a + 1024]) for a in range(0, 24 * 256, 1024)], axis=2)
    yn = np.concatenate([narrow.process(np.ascontiguousarray(x[:, :6, a:a + 1024])) for a in range(0, 24 * 256, 1024)], axis=2)
    assert np.array_equal(yw, yn)
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), FS, oracle.InputLayout.surround51)
    ref = oracle.direct_conv_f64(x[3, :6], h)
    assert np.abs(yw[3] - ref).max() <= MAX_ABS and snr_db(ref, yw[3]) >= SNR_DB


def test_long_run_does_not_drift(aw, hrtf_path):
    """600 blocks (3.2 s of audio, the FDL ring wraps 35 times, ragged callback sizes in the middle): the output stays within the
    bound of float64 convolution from the first sample to the last."""
    from scipy.signal import fftconvolve
    lay = aw.InputLayout.surround71()
    wav_o = oracle.load_wav(hrtf_path("RoomSH1.0"))
    h = oracle.hrir_matrix(wav_o, FS, oracle.InputLayout.surround71)
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), FS, lay, 256)
    n, frames = 6, 600 * 256
    x = oracle.synth_block(SEED, [2, 9, 4, 7, 1, 8], 8, 0, frames)
    eng = aw.BinauralEngine(n, 8, 256, FS, max_frames_per_call=4096)
    eng.set_bank(bank)
    sizes, pos, outs = [], 0, []
    rng = np.random.default_rng(5)
    while pos < frames:
        size = 4096 if (pos // 4096) % 7 != 3 else int(rng.integers(1, 4096))   # mostly aligned, now and then ragged
        size = min(size, frames - pos)
        outs.append(eng.process(np.ascontiguousarray(x[:, :, pos:pos + size])))
        pos += size
    y = np.concatenate(outs, axis=2)
    eng.close()
    # ragged calls switch the engine to the adapter (pending/FIFO) path for good: from the first ragged call on the output is the
    # convolution delayed by a whole number of frames d < 256 (RealtimeAudioProcessor.swift:88-116); find d from stream 0
    ref0 = np.zeros((2, frames))
    for s in range(8):
        ref0[0] += fftconvolve(x[0, s].astype(np.float64), h[s, 0].astype(np.float64))[:frames]
        ref0[1] += fftconvolve(x[0, s].astype(np.float64), h[s, 1].astype(np.float64))[:frames]
    tail = slice(frames - 20000, frames)
    best = min(range(0, 257), key=lambda d: np.abs(y[0, 0, tail] - np.roll(ref0[0], d)[tail]).max())
    for i in range(n):
        ref = np.zeros((2, frames))
        for s in range(8):
            ref[0] += fftconvolve(x[i, s].astype(np.float64), h[s, 0].astype(np.float64))[:frames]
            ref[1] += fftconvolve(x[i, s].astype(np.float64), h[s, 1].astype(np.float64))[:frames]
        got, want = y[i][:, tail], np.roll(ref, best, axis=1)[:, tail]
        assert np.abs(got - want).max() <= MAX_ABS, (i, best)
        assert snr_db(want, got) >= SNR_DB
        assert np.abs(y[i][:, :4096] - ref[:, :4096]).max() <= MAX_ABS       # before the first ragged call: no delay at all


def test_short_responses_with_many_blocks_per_call(aw, hrtf_path):
    """More blocks in a call than the ring has slots (P = 1, 2, 3 with calls of up to 64 blocks): every slot of a (stream, speaker)
    is rewritten several times inside one launch, the forward transform always one item ahead."""
    wav_o = oracle.load_wav(hrtf_path("NeutralSH1.0"))
    lay = aw.InputLayout.stereo()
    l, r = _maps(aw, lay)
    for taps, block, per_call_blocks, n in [(50, 64, 64, 9), (100, 64, 64, 150), (130, 64, 16, 9), (200, 256, 16, 33), (600, 256, 8, 5), (700, 512, 8, 9)]:
        short = wav_o.audioData[:, :taps].copy()
        bank = aw.HRIRBank(short, FS, FS, l, r, block)
        assert bank.partitions == -(-taps // block)
        blocks = 2 * per_call_blocks
        xu, y, plan = _render_twins(aw, bank, n, 2, block, blocks=blocks, unique=min(n, 9), per_call=per_call_blocks * block, pcm_lr=None)
        assert plan["kernels"][0].startswith("k_persistent<"), plan
        hs = np.stack([np.stack([short[l[s]], short[r[s]]]) for s in range(2)])
        for i in range(min(n, 9)):
            ref = oracle.direct_conv_f64(xu[i], hs)
            assert np.abs(y[i] - ref).max() <= MAX_ABS, (taps, block, per_call_blocks, i)
        for i in range(9, n):
            assert np.array_equal(y[i], y[i % 9]), (taps, block, i)


def test_speakers_sharing_a_filter_pair_share_one_delay_line(aw, hrtf_path):
    """FC and LFE use the same HRIR pair in the HeSuVi maps (VirtualSpeaker.swift:281-283): the block kernel adds the two input
    channels before the forward transform and keeps one frequency-domain delay line for both (conv(a,h) + conv(b,h) =
    conv(a+b,h)).  Same samples as one delay line per speaker up to float32 rounding, both within the oracle bound, for single- and
    multi-block calls (ragged calls, where the overlap buffer holds the summed block, are compared with the reference adapter
    sample by sample in test_arbitrary_frame_counts_follow_the_reference_adapter)."""
    wav = aw.WAVLoader.load(hrtf_path("RoomSH1.0"))
    for layout, S, rows in (("surround71", 8, 7), ("surround51", 6, 5), ("stereo", 2, 2)):
        assert aw.HRIRBank.from_wav(wav, FS, getattr(aw.InputLayout, layout)(), 256).rows == rows
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), FS, oracle.InputLayout.surround71)
    for block, n, per_call in ((256, 37, 1024), (64, 150, 640), (1024, 9, 1024)):
        bank = aw.HRIRBank.from_wav(wav, FS, aw.InputLayout.surround71(), block)
        blocks = 24
        xu, merged, plan = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=4, per_call=per_call, pcm_lr=None)
        _, separate, _ = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=4, per_call=per_call, pcm_lr=None,
                                       env={"AW_KP_MERGE_ROWS": "0"})
        assert plan["kernels"][0].startswith("k_persistent<")
        assert not np.array_equal(merged, separate) and np.abs(merged - separate).max() <= 2e-6
        for y in (merged, separate):
            for i in range(4):
                ref = oracle.direct_conv_f64(xu[i], h)
                assert np.abs(y[i] - ref).max() <= MAX_ABS and snr_db(ref, y[i]) >= SNR_DB
            for i in range(4, n):
                assert np.array_equal(y[i], y[i % 4])
    # the other execution paths keep the same rows: the three-kernel path (the only one for B < 64 and B = 4096) and the one-launch
    # fused kernel sum the channels of a row in their forward stage and walk the bank's compacted rows
    for block, n, per_call, env, first in ((256, 21, 512, {"AW_FUSED_TILE": "0"}, "k_input_rfft<"), (256, 21, 768, {"AW_PERSISTENT": "0"}, "k_fused<"),
                                           (32, 40, 96, {}, "k_input_rfft<"), (4096, 5, 4096, {}, "k_input_rfft<")):
        bank = aw.HRIRBank.from_wav(wav, FS, aw.InputLayout.surround71(), block)
        blocks = 12 if block < 4096 else 6
        xu, merged, plan = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=4, per_call=per_call, pcm_lr=None, env=env)
        _, separate, _ = _render_twins(aw, bank, n, 8, block, blocks=blocks, unique=4, per_call=per_call, pcm_lr=None,
                                       env=dict(env, AW_KP_MERGE_ROWS="0"))
        assert plan["kernels"][0].startswith(first), plan["kernels"]
        assert not np.array_equal(merged, separate) and np.abs(merged - separate).max() <= 2e-6
        for y in (merged, separate):
            for i in range(4):
                ref = oracle.direct_conv_f64(xu[i], h)
                assert np.abs(y[i] - ref).max() <= MAX_ABS and snr_db(ref, y[i]) >= SNR_DB, (block, env, i)
            for i in range(4, n):
                assert np.array_equal(y[i], y[i % 4]), (block, env, i)
