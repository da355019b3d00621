"""GPU parity tests of the convolution path, through the C ABI (airwave_b200 -> libairwave_cuda.so).

Ports of AirwaveTests/ConvolutionEngineTests.swift and AirwaveTests/RealtimeAudioProcessorTests.swift run
against the CUDA path; then CUDA vs the CPU oracle (restatement of the reference) and vs float64 direct
convolution on the bundled HRIR presets.  Tolerance (BASELINE.json north_star): float32 output within
max-abs 1e-5 and SNR >= 100 dB of the float64 oracle."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, snr_db

pytestmark = pytest.mark.gpu

KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))
MAX_ABS = 1e-5
SNR_DB = 100.0
SEED = 0x41495257


@pytest.fixture(scope="module")
def aw():
    import airwave_b200
    assert airwave_b200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return airwave_b200


# ---- ConvolutionEngineTests.swift:12-59 --------------------------------------------------------
def test_impulse_preserves_sample_order(aw):
    k = KAT["convolution_identity"]
    e = aw.ConvolutionEngine(k["hrir"], k["blockSize"])
    out = e.process(k["input"])
    assert np.all(np.abs(out - np.asarray(k["input"], np.float32)) < k["tolerance"])
    assert np.abs(out - np.asarray(k["input"], np.float32)).max() < 1e-6


def test_reset_clears_overlap_and_frequency_history(aw):
    e = aw.ConvolutionEngine([1, 0, 0, 0, 0, 0, 0, 0], 8)
    x = np.zeros(8, np.float32)
    x[-1] = 1
    e.process(x)
    e.reset()
    assert np.all(np.abs(e.process(np.zeros(8, np.float32))) < 1e-4)


def test_multiple_blocks_remain_finite_and_match_oracle(aw):
    e = aw.ConvolutionEngine([1, 0, 0, 0, 0, 0, 0, 0], 8)
    o = oracle.ConvolutionEngine([1, 0, 0, 0, 0, 0, 0, 0], 8)
    x = (np.arange(8) / 7).astype(np.float32)
    for _ in range(64):
        y = e.process(x)
        assert np.all(np.isfinite(y)) and np.abs(y - o.process(x)).max() < 1e-6
        x = (-x * np.float32(0.97) + np.float32(0.01)).astype(np.float32)


def test_identical_input_after_reset_produces_identical_output(aw):
    e = aw.ConvolutionEngine([1, 0, 0, 0, 0, 0, 0, 0], 8)
    x = np.arange(-0.75, 0.7501, 0.2, dtype=np.float32)[:8]
    a = e.process(x)
    e.reset()
    assert np.array_equal(a, e.process(x))


def test_wrong_frame_count_is_a_no_op_and_bad_block_size_is_rejected(aw):
    assert aw.ConvolutionEngine([1.0], 8).process(np.zeros(8, np.float32), frameCount=7) is None
    with pytest.raises(ValueError):
        aw.ConvolutionEngine([1.0], 12)


def test_process_and_accumulate(aw):
    e = aw.ConvolutionEngine([0.5, 0.25], 8)
    acc = np.ones(8, np.float32)
    x = np.arange(8, dtype=np.float32)
    e.processAndAccumulate(x, acc)
    want = 1 + 0.5 * x + 0.25 * np.concatenate([[0], x[:-1]])
    assert np.abs(acc - want).max() < 1e-5


# ---- RealtimeAudioProcessorTests.swift:59-126 ---------------------------------------------------
def make_processor(aw, renderer_count=2, block=512, max_frames=4096):
    renderers = [aw.VirtualSpeakerRenderer("FL" if i == 0 else "FR", aw.ConvolutionEngine([float(i + 1)], block),
                                           aw.ConvolutionEngine([float(i + 1)], block)) for i in range(renderer_count)]
    return aw.RealtimeAudioProcessor(renderers, block, max_frames)


def run(p, size, left=1.0, right=2.0):
    return p.process(np.full(size, left, np.float32), np.full(size, right, np.float32))


@pytest.mark.parametrize("size", KAT["adapter_sizes"]["sizes"])
def test_all_required_callback_sizes_write_finite_output(aw, size):
    l, r = run(make_processor(aw), size)
    assert l.shape == (size,) and np.all(np.isfinite(l)) and np.all(np.isfinite(r))


def test_mixed_callback_sequence_preserves_order_after_adapter_latency(aw):
    k = KAT["adapter_mixed_sequence"]
    p = make_processor(aw, 1)
    out = np.concatenate([run(p, n)[0] for n in k["sizes"]])
    assert len(out) == k["total"]
    assert np.all(out[:k["leading_zeros"]] == 0)
    assert np.all(np.abs(out[k["leading_zeros"]:] - k["then_value"]) < k["tolerance"])


def test_reset_clears_pending_input_and_queued_output(aw):
    p = make_processor(aw, 1)
    run(p, 512)
    p.reset()
    l, r = run(p, 1)
    assert l.tolist() == [0] and r.tolist() == [0]


def test_underflow_silence_and_mono_duplication(aw):
    p = make_processor(aw, 1)
    l, r = run(p, 3, 0.5, 0.5)
    assert l.tolist() == [0, 0, 0] and r.tolist() == [0, 0, 0]
    l, r = run(p, 512, 0.5, 0.5)
    assert np.array_equal(l, r)


def test_canaries_aliased_outputs_and_nil_right_input(aw):
    import ctypes as C
    p = make_processor(aw, 1)
    size, canary = 4096, np.float32(12345)
    inp = np.zeros(size + 2, np.float32)
    outp = np.full(size + 2, canary, np.float32)
    eng = p._engine
    rc = aw.lib().aw_engine_process_stereo(eng._h, inp[1:].ctypes.data_as(C.c_void_p), None, outp[1:].ctypes.data_as(C.c_void_p),
                                           outp[1:].ctypes.data_as(C.c_void_p), size)
    assert rc == 0
    assert inp[0] == 0 and inp[-1] == 0 and outp[0] == canary and outp[-1] == canary
    assert np.all(np.isfinite(outp))


def test_frame_count_above_maximum_is_a_precondition_failure(aw):
    with pytest.raises(AssertionError):
        run(make_processor(aw, 1, 512, 1024), 1025)


def test_literal_stereo_two_renderers_matches_oracle_on_arbitrary_callback_sizes(aw, hrtf_path):
    wav_o = oracle.load_wav(hrtf_path("NeutralSH1.0"))
    ro = oracle.activate_preset(wav_o, 48000.0, oracle.InputLayout.stereo, 512)
    po = oracle.RealtimeAudioProcessor(ro, 512, 4096, literalStereo=True)
    rg = [aw.VirtualSpeakerRenderer(r.speaker, aw.ConvolutionEngine(r.convolverLeftEar.hrirSamples, 512),
                                    aw.ConvolutionEngine(r.convolverRightEar.hrirSamples, 512)) for r in ro]
    pg = aw.RealtimeAudioProcessor(rg, 512, 4096)
    rng = np.random.default_rng(7)
    pos = 0
    for size in [1, 64, 128, 511, 512, 513, 768, 1024, 4096, 37, 2048, 512, 512, 999]:
        x = oracle.synth_block(SEED, [5], 2, pos, size)[0]
        pos += size
        mono = size == 37
        gl, gr = pg.process(x[0], None if mono else x[1])
        ol, orr = po.process(x[0], None if mono else x[1])
        assert np.abs(gl - ol).max() <= 2e-6 and np.abs(gr - orr).max() <= 2e-6, size


# ---- CUDA vs oracle vs float64 direct convolution on the bundled presets --------------------------
CASES = [  # (preset, layout, block, streams, blocks)  — C1, C2 and the block-size sweep of C5 at test scale
    ("NeutralSH1.0", "stereo", 512, 1, 24),
    ("RoomSH1.0", "surround71", 256, 6, 40),
    ("StageSH1.0", "surround71", 64, 3, 150),
    ("RoomSH1.0", "surround71", 128, 2, 80),
    ("RoomSH1.0", "surround71", 1024, 2, 12),
    ("StageSH1.0", "surround71", 2048, 2, 6),
    ("NeutralSH1.0", "surround71", 4096, 2, 4),
    ("RoomSH1.0", "surround51", 512, 3, 20),
]


@pytest.mark.parametrize("preset,layout,block,streams,blocks", CASES)
def test_render_matches_oracle_and_float64_direct_convolution(aw, hrtf_path, preset, layout, block, streams, blocks):
    wav = aw.WAVLoader.load(hrtf_path(preset))
    lay = getattr(aw.InputLayout, layout)()
    bank = aw.HRIRBank.from_wav(wav, 48000.0, lay, block)
    S = len(lay.channels)
    assert (bank.n_speakers, bank.block, bank.partitions, bank.taps) == (S, block, -(-4320 // block), 4320)
    eng = aw.BinauralEngine(streams, S, block, 48000.0, max_frames_per_call=min(4096, 4 * block))
    eng.set_bank(bank)
    ids = [11 * i + 1 for i in range(streams)]
    frames = blocks * block
    x = oracle.synth_block(SEED, ids, S, 0, frames)
    per_call = min(4096, 4 * block)
    got = np.concatenate([eng.process(x[:, :, a:a + per_call]) for a in range(0, frames, per_call)], axis=2)
    wav_o = oracle.load_wav(hrtf_path(preset))
    lay_o = getattr(oracle.InputLayout, layout)
    h = oracle.hrir_matrix(wav_o, 48000.0, lay_o)
    for i in range(streams):
        ref64 = oracle.direct_conv_f64(x[i], h)
        assert np.abs(got[i] - ref64).max() <= MAX_ABS
        assert snr_db(ref64, got[i]) >= SNR_DB
    # the CPU restatement of the reference, stream 0
    rap = oracle.RealtimeAudioProcessor(oracle.activate_preset(wav_o, 48000.0, lay_o, block), block, block, literalStereo=False)
    ref32 = np.concatenate([np.stack(rap.process_channels([x[0, s, b * block:(b + 1) * block] for s in range(S)]))
                            for b in range(blocks)], axis=1)
    assert np.abs(got[0] - ref32).max() <= MAX_ABS
    assert snr_db(ref32, got[0]) >= SNR_DB


def test_long_brir_65536_taps_wraps_the_ring(aw):
    """C3 at test scale: P = 128 partitions, more than P blocks so the FDL ring wraps (Q4)."""
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(1)
    taps, block, S = 65536, 512, 8
    n = np.arange(taps)
    pcm = (0.05 * rng.standard_normal((14, taps)) * np.exp(-n / (0.25 * 48000.0))).astype(np.float32)
    pcm[:, 190] += 0.5
    lay = aw.InputLayout.surround71()
    m = aw.HRIRChannelMap.hesuvi14Channel(lay.channels)
    l = [m.getIndices(s)[0] for s in lay.channels]
    r = [m.getIndices(s)[1] for s in lay.channels]
    bank = aw.HRIRBank(pcm, 48000.0, 48000.0, l, r, block)
    assert bank.partitions == 128
    eng = aw.BinauralEngine(2, S, block, 48000.0, 4096)
    eng.set_bank(bank)
    blocks = 136
    x = oracle.synth_block(SEED, [0, 9], S, 0, blocks * block)
    got = np.concatenate([eng.process(x[:, :, a:a + 4096]) for a in range(0, blocks * block, 4096)], axis=2)
    for i in range(2):
        ref = np.zeros((2, blocks * block))
        for s in range(S):
            ref[0] += fftconvolve(x[i, s].astype(np.float64), pcm[l[s]].astype(np.float64))[: blocks * block]
            ref[1] += fftconvolve(x[i, s].astype(np.float64), pcm[r[s]].astype(np.float64))[: blocks * block]
        assert np.abs(got[i] - ref).max() <= MAX_ABS
        assert snr_db(ref, got[i]) >= SNR_DB


def test_filter_bank_matches_float64_fft_of_partitions(aw, hrtf_path):
    """K1 alone: bank[s][p][k] = rfft(h[pB:(p+1)B] || 0_B)[k] * 2 * 0.25/N (ConvolutionEngine.swift:143-182, :356)."""
    wav = aw.WAVLoader.load(hrtf_path("RoomSH1.0"))
    block = 256
    bank = aw.HRIRBank.from_wav(wav, 48000.0, aw.InputLayout.surround71(), block)
    spec, ny = bank.read()
    h = oracle.hrir_matrix(oracle.load_wav(hrtf_path("RoomSH1.0")), 48000.0, oracle.InputLayout.surround71)
    P = bank.partitions
    hp = np.zeros((8, 2, P * block), np.float64)
    hp[:, :, :4320] = h
    X = np.fft.rfft(np.concatenate([hp.reshape(8, 2, P, block), np.zeros((8, 2, P, block))], axis=3), axis=3) * (0.5 / (2 * block))
    got = spec[..., 0::2] + 1j * spec[..., 1::2]          # [s][p][k][ear]
    want = np.moveaxis(X[..., :block], 1, 3)               # [s][p][k][ear]
    scale = np.abs(want).max()
    assert np.abs(got[:, :, 1:] - want[:, :, 1:]).max() <= 2e-6 * scale
    assert np.abs(got[:, :, 0].real - want[:, :, 0].real).max() <= 2e-6 * scale and np.all(got[:, :, 0].imag == 0)
    assert np.abs(ny - np.moveaxis(X[..., block].real, 1, 2)).max() <= 2e-6 * scale


def test_resampler_is_bit_exact_with_the_oracle_and_feeds_the_bank(aw, hrtf_path):
    wav_o = oracle.load_wav(hrtf_path("StageSH1.0"))
    x = wav_o.audioData[3]
    got = aw.Resampler.resampleHighQuality(x, 44100.0, 48000.0)
    want = oracle.resample_high_quality(x, 44100.0, 48000.0)
    assert len(got) == 4702 and np.array_equal(got, want)
    assert np.array_equal(aw.Resampler.resampleHighQuality(x, 48000.0, 48000.004), x)
    with pytest.raises(aw.AirwaveError) as e:
        aw.Resampler.resampleHighQuality(x, 48000.0, 44100.0)
    assert e.value.status == 11
    # C4's resample step: the preset relabelled as a 44.1 kHz source: 4320 -> 4702 taps, P = 19 at B = 256
    m = aw.HRIRChannelMap.hesuvi14Channel(aw.InputLayout.surround71().channels)
    sp = aw.InputLayout.surround71().channels
    bank = aw.HRIRBank(wav_o.audioData, 44100.0, 48000.0, [m.getIndices(s)[0] for s in sp], [m.getIndices(s)[1] for s in sp], 256)
    assert (bank.taps, bank.partitions) == (4702, 19)
    wav_o.sampleRate = 44100.0
    h = oracle.hrir_matrix(wav_o, 48000.0, oracle.InputLayout.surround71)
    eng = aw.BinauralEngine(1, 8, 256, 48000.0, 1024)
    eng.set_bank(bank)
    xin = oracle.synth_block(SEED, [2], 8, 0, 30 * 256)
    got = np.concatenate([eng.process(xin[:, :, a:a + 1024]) for a in range(0, 30 * 256, 1024)], axis=2)[0]
    ref = oracle.direct_conv_f64(xin[0], h)
    assert np.abs(got - ref).max() <= MAX_ABS and snr_db(ref, got) >= SNR_DB


def test_flagged_correct_resampler_mode(aw, hrtf_path):
    """AW_RESAMPLE_CORRECT (SURVEY.md Q7's flagged alternative to the literal Resampler.swift:31-68): time-correct linear
    interpolation, bit-exact with the float64 numpy restatement, up- and down-sampling; the literal mode stays the default."""
    wav_o = oracle.load_wav(hrtf_path("StageSH1.0"))
    x = wav_o.audioData[3]
    for src, dst in [(44100.0, 48000.0), (96000.0, 48000.0), (48000.0, 44100.0), (88200.0, 48000.0)]:
        got = aw.Resampler.resampleHighQuality(x, src, dst, correct=True)
        want = oracle.resample_linear_f64(x, src, dst)
        assert len(got) == int(len(x) / (src / dst)) and np.array_equal(got, want), (src, dst)
    # a slow sinusoid sampled at 44.1 kHz comes out as the same sinusoid sampled at 48 kHz (the literal mode plays it 8.8 % fast)
    t = np.arange(4410) / 44100.0
    tone = np.sin(2 * np.pi * 200.0 * t).astype(np.float32)
    up = aw.Resampler.resampleHighQuality(tone, 44100.0, 48000.0, correct=True)
    want = np.sin(2 * np.pi * 200.0 * np.arange(len(up)) / 48000.0)
    assert np.abs(up[:-2] - want[:-2]).max() < 2e-4
    literal = aw.Resampler.resampleHighQuality(tone, 44100.0, 48000.0)
    assert np.abs(literal[:4000] - want[:4000]).max() > 0.5
    # down-sampled bank: 96 kHz -> 48 kHz HRIR (refused in the literal mode), rendered against float64 convolution with the same taps
    m = aw.HRIRChannelMap.hesuvi14Channel(aw.InputLayout.surround71().channels)
    sp = aw.InputLayout.surround71().channels
    l, r = [m.getIndices(s)[0] for s in sp], [m.getIndices(s)[1] for s in sp]
    with pytest.raises(aw.AirwaveError) as e:
        aw.HRIRBank(wav_o.audioData, 96000.0, 48000.0, l, r, 256)
    assert e.value.status == 11
    bank = aw.HRIRBank(wav_o.audioData, 96000.0, 48000.0, l, r, 256, correct_resampling=True)
    assert (bank.taps, bank.partitions) == (2160, 9)
    h = np.stack([np.stack([oracle.resample_linear_f64(wav_o.audioData[l[s]], 96000.0, 48000.0),
                            oracle.resample_linear_f64(wav_o.audioData[r[s]], 96000.0, 48000.0)]) for s in range(8)])
    eng = aw.BinauralEngine(1, 8, 256, 48000.0, 1024)
    eng.set_bank(bank)
    xin = oracle.synth_block(SEED, [5], 8, 0, 16 * 256)
    got = np.concatenate([eng.process(xin[:, :, a:a + 1024]) for a in range(0, 16 * 256, 1024)], axis=2)[0]
    ref = oracle.direct_conv_f64(xin[0], h)
    assert np.abs(got - ref).max() <= MAX_ABS and snr_db(ref, got) >= SNR_DB


def test_bank_errors_mirror_the_reference(aw):
    pcm = np.ones((7, 16), np.float32)
    with pytest.raises(aw.AirwaveError) as e:
        aw.HRIRBank(pcm, 48000.0, 48000.0, [0, 8], [1, 7], 8)       # HRIRManager.swift:375-379
    assert e.value.status == 6 and "out of range for 7 channels" in e.value.message
    with pytest.raises(aw.AirwaveError) as e:
        aw.HRIRBank(pcm, 48000.0, 48000.0, [-1], [-1], 8)           # :420-422
    assert e.value.status == 7
    with pytest.raises(aw.AirwaveError) as e:
        aw.HRIRBank(pcm, 48000.0, 48000.0, [0], [1], 12)
    assert e.value.status == 4
    seven = aw.HRIRChannelMap.hesuvi7Channel(aw.InputLayout.surround71().channels)
    assert seven.getIndices("LFE") == (2, 2)


def test_arbitrary_frame_counts_follow_the_reference_adapter(aw, hrtf_path):
    """K7: ragged callback sizes go through pending/FIFO exactly like RealtimeAudioProcessor.swift:77-190 —
    silence on underflow, growing latency — and carry the same samples as block-aligned rendering."""
    wav = aw.WAVLoader.load(hrtf_path("RoomSH1.0"))
    bank = aw.HRIRBank.from_wav(wav, 48000.0, aw.InputLayout.surround71(), 256)
    x = oracle.synth_block(SEED, [0, 1, 2], 8, 0, 8192)
    a = aw.BinauralEngine(3, 8, 256, 48000.0, 4096)
    a.set_bank(bank)
    aligned = np.concatenate([a.process(x[:, :, i:i + 1024]) for i in range(0, 8192, 1024)], axis=2)
    b = aw.BinauralEngine(3, 8, 256, 48000.0, 4096)
    b.set_bank(bank)
    wav_o = oracle.load_wav(hrtf_path("RoomSH1.0"))
    raps = [oracle.RealtimeAudioProcessor(oracle.activate_preset(wav_o, 48000.0, oracle.InputLayout.surround71, 256), 256, 4096,
                                          literalStereo=False) for _ in range(3)]
    sizes, pos = [1, 255, 256, 257, 100, 4096, 33, 512, 1000, 1682], 0
    zeros_seen = 0
    for n in sizes:
        got = b.process(x[:, :, pos:pos + n])
        for i in range(3):
            ol, orr = raps[i].process_channels([x[i, s, pos:pos + n] for s in range(8)])
            assert np.abs(got[i, 0] - ol).max() <= 2e-6 and np.abs(got[i, 1] - orr).max() <= 2e-6, (n, i)
            assert np.array_equal(got[i, 0] == 0, ol == 0)     # silence exactly where the reference underflows
        zeros_seen += int((got[0, 0] == 0).sum())
        pos += n
    assert pos == 8192 and zeros_seen >= 1 + 100
    # the adapter only delays: block k of the ragged run is bit-identical to block k of the aligned run
    c = aw.BinauralEngine(3, 8, 256, 48000.0, 4096)
    c.set_bank(bank)
    first = c.process(x[:, :, :128])
    assert np.all(first == 0)
    rest = np.concatenate([c.process(x[:, :, 128 + i:128 + i + 1024]) for i in range(0, 4096, 1024)], axis=2)
    assert np.array_equal(rest, aligned[:, :, :4096])   # 128 frames late, otherwise the very same samples


def test_streams_are_independent_and_batch_size_does_not_change_results(aw, hrtf_path):
    """Row (e): sharding by stream must give bit-identical per-stream output for any batch/tile size."""
    wav = aw.WAVLoader.load(hrtf_path("RoomSH1.0"))
    bank = aw.HRIRBank.from_wav(wav, 48000.0, aw.InputLayout.surround71(), 256)
    n = 37
    x = oracle.synth_block(SEED, range(n), 8, 0, 2048)
    big = aw.BinauralEngine(n, 8, 256, 48000.0, 2048)
    big.set_bank(bank)
    y = big.process(x)
    for first, count in [(0, 1), (5, 9), (20, 17)]:
        small = aw.BinauralEngine(count, 8, 256, 48000.0, 2048)
        small.set_bank(bank)
        assert np.array_equal(small.process(x[first:first + count]), y[first:first + count])


def test_per_range_banks_passthrough_and_reset(aw, hrtf_path):
    room = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), 48000.0, aw.InputLayout.stereo(), 256)
    stage = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("StageSH1.0")), 48000.0, aw.InputLayout.stereo(), 256)
    eng = aw.BinauralEngine(6, 2, 256, 48000.0, 1024, max_partitions=17)
    x = oracle.synth_block(SEED, range(6), 2, 0, 1024)
    assert np.array_equal(eng.process(x), x)                     # no renderers: passthrough (HRIRManager.swift:555-564)
    eng.set_bank(room, 0, 2)
    eng.set_bank(stage, 2, 2)                                    # streams 4, 5 stay passthrough
    y = eng.process(x)
    assert np.array_equal(y[4:], x[4:])
    for bank, sl in [(room, slice(0, 2)), (stage, slice(2, 4))]:
        ref = aw.BinauralEngine(2, 2, 256, 48000.0, 1024)
        ref.set_bank(bank)
        assert np.array_equal(ref.process(x[sl]), y[sl])
    y2 = eng.process(x)
    eng.reset(0, 4)
    assert np.array_equal(eng.process(x)[:4], y[:4]) and not np.array_equal(y2[:4], y[:4])
    with pytest.raises(aw.AirwaveError) as e:
        eng.set_bank(room, 4, 5)
    assert e.value.status == 8


def test_fft_plan_cache(aw):
    aw.FFTSetupManager.getSetup(9)
    aw.FFTSetupManager.getSetup(9)
    count, sizes = aw.FFTSetupManager.getCacheStats()
    assert 512 in sizes and count == len(sizes) == len(set(sizes))


def test_device_synthetic_input_is_the_oracles_generator(aw):
    """bench.py fills its device-resident input with aw_synth_fill_device and times the CPU baseline on oracle.synth_fill
    (counter-based generator keyed (seed, stream, speaker, frame)): both must be the same data, bit for bit."""
    torch = pytest.importorskip("torch")
    n, S, frames, first, frame0 = 5, 8, 777, 1234, 100000
    buf = torch.empty((n, S, frames), dtype=torch.float32, device="cuda:0")
    aw._lib.check(aw.lib().aw_synth_fill_device(0, buf.data_ptr(), first, n, S, frame0, frames, SEED, None))
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    want = oracle.synth_block(SEED, range(first, first + n), S, frame0, frames)
    assert np.array_equal(got, want)
    assert np.abs(got).max() <= 0.25 and got.std() > 0.1      # uniform in [-0.25, 0.25]
