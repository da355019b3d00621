"""Pins the CPU oracle against the reference's ConvolutionEngineTests / RealtimeAudioProcessorTests
(AirwaveTests/ConvolutionEngineTests.swift, AirwaveTests/RealtimeAudioProcessorTests.swift) and
against float64 direct convolution on the bundled HRIR presets."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, snr_db

KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))
BLOCK = 8


def make_engine():
    return oracle.ConvolutionEngine(KAT["convolution_identity"]["hrir"], BLOCK)


# --- ConvolutionEngineTests.swift:12-59 ---------------------------------------------------------
def test_impulse_preserves_sample_order():
    k = KAT["convolution_identity"]
    out = make_engine().process(k["input"])
    assert np.all(np.abs(out - np.asarray(k["input"], np.float32)) < k["tolerance"])


def test_reset_clears_overlap_and_frequency_history():
    e = make_engine()
    x = np.zeros(BLOCK, np.float32)
    x[-1] = 1
    e.process(x)
    e.reset()
    out = e.process(np.zeros(BLOCK, np.float32))
    assert np.all(np.abs(out) < 1e-4)


def test_multiple_blocks_remain_finite():
    e = make_engine()
    x = (np.arange(BLOCK) / 7).astype(np.float32)
    for _ in range(64):
        assert np.all(np.isfinite(e.process(x)))
        x = (-x * np.float32(0.97) + np.float32(0.01)).astype(np.float32)


def test_identical_input_after_reset_produces_identical_output():
    e = make_engine()
    x = np.arange(-0.75, 0.7501, 0.2, dtype=np.float32)[:BLOCK]
    a = e.process(x)
    e.reset()
    b = e.process(x)
    assert np.all(np.abs(a - b) < 1e-4)


def test_wrong_frame_count_is_a_no_op():
    # ConvolutionEngine.swift:370-372
    assert make_engine().process(np.zeros(BLOCK, np.float32), frameCount=BLOCK - 1) is None


def test_non_power_of_two_block_is_rejected():
    with pytest.raises(ValueError):
        oracle.ConvolutionEngine([1.0], 12)


# --- zrip conventions -----------------------------------------------------------
@pytest.mark.parametrize("log2n", [4, 7, 10, 13])
def test_zrip_forward_is_twice_the_dft_in_packed_layout(log2n):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    re, im = x[0::2].copy(), x[1::2].copy()
    s = oracle.FFTSetup(log2n)
    s.zrip(re, im, True)
    X = 2 * np.fft.rfft(x.astype(np.float64))
    scale = np.abs(X).max()
    assert abs(re[0] - X[0].real) < 1e-5 * scale and abs(im[0] - X[n // 2].real) < 1e-5 * scale
    got = re[1:] + 1j * im[1:]
    assert np.abs(got - X[1:n // 2]).max() < 1e-5 * scale
    # inverse is unnormalised: inverse(forward(x)) = 2N x
    s.zrip(re, im, False)
    back = np.empty(n, np.float32)
    back[0::2], back[1::2] = re, im
    assert np.abs(back / (2 * n) - x).max() < 1e-5


# --- multi-partition + real HRIR vs float64 direct convolution (what the reference never tests) --
@pytest.mark.parametrize("preset,layout,block", [("NeutralSH1.0", "stereo", 512), ("RoomSH1.0", "surround71", 256),
                                                 ("StageSH1.0", "surround71", 64)])
def test_oracle_upols_matches_float64_direct_convolution(preset, layout, block, hrtf_path):
    wav = oracle.load_wav(hrtf_path(preset))
    lay = getattr(oracle.InputLayout, layout)
    renderers = oracle.activate_preset(wav, 48000.0, lay, block)
    S = len(renderers)
    assert renderers[0].convolverLeftEar.partitionCount == -(-4320 // block)
    rap = oracle.RealtimeAudioProcessor(renderers, block, block, literalStereo=False)
    blocks = 40 if block >= 256 else 90
    x = oracle.synth_block(0x41495257, [3], S, 0, blocks * block)[0]
    outL, outR = [], []
    for b in range(blocks):
        l, r = rap.process_channels([x[s, b * block:(b + 1) * block] for s in range(S)])
        outL.append(l)
        outR.append(r)
    got = np.stack([np.concatenate(outL), np.concatenate(outR)])
    ref = oracle.direct_conv_f64(x, oracle.hrir_matrix(wav, 48000.0, lay))
    assert np.abs(got - ref).max() <= 1e-5
    assert snr_db(ref, got) >= 100.0


def test_literal_stereo_uses_at_most_two_renderers(hrtf_path):
    # RealtimeAudioProcessor.swift:145 — min(renderers.count, 2)
    wav = oracle.load_wav(hrtf_path("RoomSH1.0"))
    r8 = oracle.activate_preset(wav, 48000.0, oracle.InputLayout.surround71, 64)
    r2 = oracle.activate_preset(wav, 48000.0, oracle.InputLayout.stereo, 64)
    x = oracle.synth_block(1, [0], 2, 0, 64 * 6)[0]
    a = oracle.RealtimeAudioProcessor(r8, 64, 4096, literalStereo=True).process(x[0], x[1])
    b = oracle.RealtimeAudioProcessor(r2, 64, 4096, literalStereo=True).process(x[0], x[1])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# --- RealtimeAudioProcessorTests.swift:59-126 ----------------------------------------------------
def make_processor(renderer_count=2, block=512, max_frames=4096):
    renderers = [oracle.VirtualSpeakerRenderer("FL" if i == 0 else "FR",
                                               oracle.ConvolutionEngine([float(i + 1)], block),
                                               oracle.ConvolutionEngine([float(i + 1)], block))
                 for i in range(renderer_count)]
    return oracle.RealtimeAudioProcessor(renderers, block, max_frames, literalStereo=True)


def run(p, size, left=1.0, right=2.0):
    return p.process(np.full(size, left, np.float32), np.full(size, right, np.float32))


@pytest.mark.parametrize("size", KAT["adapter_sizes"]["sizes"])
def test_all_required_callback_sizes_write_finite_output(size):
    l, r = run(make_processor(), size)
    assert np.all(np.isfinite(l)) and np.all(np.isfinite(r))


def test_mixed_callback_sequence_preserves_order_after_adapter_latency():
    k = KAT["adapter_mixed_sequence"]
    p = make_processor(1)
    out = np.concatenate([run(p, n)[0] for n in k["sizes"]])
    assert len(out) == k["total"]
    assert np.all(out[:k["leading_zeros"]] == 0)
    assert np.all(np.abs(out[k["leading_zeros"]:] - k["then_value"]) < k["tolerance"])


def test_reset_clears_pending_input_and_queued_output():
    p = make_processor(1)
    run(p, 512)
    p.reset()
    l, r = run(p, 1)
    assert l.tolist() == [0] and r.tolist() == [0]


def test_underflow_silence_and_mono_duplication():
    p = make_processor(1)
    l, r = run(p, 3, 0.5, 0.5)
    assert l.tolist() == [0, 0, 0] and r.tolist() == l.tolist()
    l, r = run(p, 512, 0.5, 0.5)
    assert np.array_equal(l, r)


def test_aliased_outputs_and_nil_right_input():
    p = make_processor(1)
    l, r = p.process_channels([np.zeros(4096, np.float32), None], aliasOutputs=True)
    assert l is r and np.all(np.isfinite(l))


def test_frame_count_above_maximum_is_a_precondition_failure():
    with pytest.raises(AssertionError):
        run(make_processor(1, 512, 1024), 1025)


# --- Resampler (parity unpinned by the reference: semantics of SURVEY.md Q7) ---------------------
def test_resampler_vgenp_semantics():
    x = np.arange(1, 4321, dtype=np.float32)
    y = oracle.resample_high_quality(x, 44100.0, 48000.0)
    assert len(y) == 4702 == oracle.resample_output_count(4320, 44100.0, 48000.0)
    stride = np.float32(44100.0 / 48000.0)
    last_bp = int(np.float32(4319) * stride)
    n = np.arange(1, last_bp + 1)
    # out[n] = lerp(input, n / stride): the IR is read faster (time-compressed), then the tail holds
    want = 1.0 + n / float(stride)
    assert np.abs(y[1:last_bp + 1] - want).max() < 7e-3
    assert y[0] == x[0] and np.all(y[last_bp + 1:] == x[-1])
    assert oracle.resample_high_quality(x, 48000.0, 48000.004) is not None
    assert len(oracle.resample_high_quality(x, 48000.0, 48000.004)) == 4320
    with pytest.raises(ValueError):
        oracle.resample_high_quality(x, 48000.0, 44100.0)


def test_flagged_correct_resampler_checker_properties():
    """oracle.resample_linear_f64 (checker of AW_RESAMPLE_CORRECT): output count as Resampler.swift:39, identity when the rates
    agree, exact on a linear ramp, endpoints held, and close to scipy's polyphase resampler on a band-limited signal."""
    import scipy.signal
    x = np.linspace(-1, 1, 441, dtype=np.float32)
    y = oracle.resample_linear_f64(x, 44100.0, 48000.0)
    assert len(y) == int(441 / (44100.0 / 48000.0)) == 480
    pos = np.arange(480) * (44100.0 / 48000.0)
    want = np.where(pos >= 440, 1.0, -1 + pos * (2 / 440))
    assert np.abs(y - want).max() < 1e-6
    assert np.array_equal(oracle.resample_linear_f64(x, 48000.0, 48000.004), x)
    assert len(oracle.resample_linear_f64(x, 96000.0, 48000.0)) == 220
    t = np.arange(4410) / 44100.0
    tone = (np.sin(2 * np.pi * 300 * t) * np.hanning(4410)).astype(np.float32)
    lin = oracle.resample_linear_f64(tone, 44100.0, 48000.0)
    poly = scipy.signal.resample_poly(tone.astype(np.float64), 160, 147)[: len(lin)]
    assert np.abs(lin - poly).max() < 2e-3
