"""Pins the CPU oracle against the reference's ParametricEqualizerProcessorTests
(AirwaveTests/ParametricEqualizerProcessorTests.swift) — the same cases, line by line."""
import json
import math
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))


def make_filter(type_, frequency, gain, q, enabled=True):
    return dict(type=type_, frequencyHz=frequency, gainDB=gain, q=q, isEnabled=enabled, sourceLine=1)


def magnitude_db(c, f, fs):
    w = 2 * math.pi * f / fs
    z = complex(math.cos(w), math.sin(w))
    num = c[0] + c[1] / z + c[2] / (z * z)
    den = 1 + c[3] / z + c[4] / (z * z)
    return 20 * math.log10(abs(num / den))


# :6-60
@pytest.mark.parametrize("case", KAT["biquad_coefficients"]["cases"])
def test_golden_coefficients(case):
    c = oracle.biquad_make(case["type"], case["gainDB"], case["frequencyHz"], case["q"], case["sampleRate"])
    assert np.abs(c - np.asarray(case["expected"])).max() <= KAT["biquad_coefficients"]["tolerance"]


# :62-85
@pytest.mark.parametrize("case", KAT["biquad_magnitude_db"]["cases"])
def test_golden_magnitude(case):
    fs = case["sampleRate"]
    c = oracle.biquad_make(case["type"], case["gainDB"], case["frequencyHz"], case["q"], fs)
    for f, want in zip([0, case["frequencyHz"], fs / 2 - 1], case["expected"]):
        assert abs(magnitude_db(c, f, fs) - want) <= 1e-9


# :87-107
def test_unity_and_preamp_only_states():
    unity = oracle.ParametricEqualizerProcessor.prepare(None, 48000)
    preamp = oracle.ParametricEqualizerProcessor.prepare(dict(preampDB=6, filters=[]), 48000)
    l = np.array([0.25, -0.5, 1], np.float32)
    r = np.array([-0.75, 0.5, 0.125], np.float32)
    ul, ur = unity.process(l, r)
    pl, pr = preamp.process(l, r)
    gain = np.float32(10.0 ** (6.0 / 20.0))
    assert np.array_equal(ul, l) and np.array_equal(ur, r)
    for i in (0, 2):
        assert abs(pl[i] - l[i] * gain) <= 1e-6 and abs(pr[i] - r[i] * gain) <= 1e-6


# :109-133
def test_known_impulse_response_preserves_cascade_order():
    k = KAT["cascade_impulse_response"]
    filters = [make_filter(f["type"], f["frequencyHz"], f["gainDB"], f["q"]) for f in k["filters"]]
    state = oracle.ParametricEqualizerProcessor.prepare(dict(filters=filters), k["sampleRate"])
    l, r = state.process(k["left_input"], [0] * 6)
    assert np.abs(l - np.asarray(k["expected_left"], np.float32)).max() <= k["tolerance"]
    assert np.all(r == 0)


# :135-152
def test_disabled_filters_excluded_and_subnormal_flushed():
    state = oracle.ParametricEqualizerProcessor.prepare(
        dict(filters=[make_filter("peaking", 1000, 12, 0.7, enabled=False)]), 48000)
    l, r = state.process([1, 0], [1, 0])
    assert l.tolist() == [1, 0] and r.tolist() == [1, 0]
    active = oracle.ParametricEqualizerProcessor.prepare(dict(filters=[make_filter("peaking", 1000, 6, 0.707)]), 48000)
    tiny = np.array([np.float32(1e-45), 0], np.float32)  # Float.leastNonzeroMagnitude
    l, _ = active.process(tiny, [0, 0])
    assert l[0] != 0 and l[1] == 0


# :154-189
def test_in_place_processing_preserves_canaries():
    state = oracle.ParametricEqualizerProcessor.prepare(dict(filters=[make_filter("highShelf", 6000, -5, 0.8)]), 48000)
    size, canary = 4096, np.float32(12345)
    left = np.full(size + 2, canary, np.float32)
    right = np.full(size + 2, canary, np.float32)
    idx = np.arange(size)
    left[1:-1] = (idx % 17).astype(np.float32) / 17
    right[1:-1] = -(idx % 13).astype(np.float32) / 13
    state.process_inplace(left[1:], right[1:], size)
    assert left[0] == canary and left[-1] == canary and right[0] == canary and right[-1] == canary
    assert np.all(np.isfinite(left)) and np.all(np.isfinite(right))


# :191-209
def test_preparation_rejects_invalid_inputs():
    P = oracle.ParametricEqualizerProcessor
    with pytest.raises(oracle.ParametricEqualizerPreparationError) as e:
        P.prepare(None, 0)
    assert e.value.name == "invalidSampleRate"
    with pytest.raises(oracle.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 24000, 1, 1)]), 48000)
    assert e.value.name == "invalidFilter" and e.value.index == 0 and e.value.filter_error == 2
    with pytest.raises(oracle.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 1000, 1, 0)]), 48000)
    assert e.value.filter_error == 3
    with pytest.raises(oracle.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 500 + i, 1, 1) for i in range(65)]), 48000)
    assert e.value.name == "tooManyFilters"


def run(p, n, lv, rv):
    return p.process(np.full(n, lv, np.float32), np.full(n, rv, np.float32))


# :211-232
@pytest.mark.parametrize("fs", [44100.0, 48000.0, 96000.0])
def test_crossfade_uses_exact_twenty_millisecond_ramp(fs):
    p = oracle.ParametricEqualizerProcessor(fs, 4096)
    gain = np.float32(10.0 ** (6.0 / 20.0))
    p.setTarget(dict(preampDB=6))
    length = max(1, int(round(fs * 0.020)))
    first_half = max(1, length // 2)
    a, _ = run(p, first_half, 1, 1)
    b, br = run(p, length - first_half, 1, 1)
    assert abs(a[0] - (1 + (gain - 1) / np.float32(length))) <= 1e-5
    assert abs(b[-1] - gain) <= 1e-5 and abs(br[-1] - gain) <= 1e-5
    assert np.all(np.isfinite(np.concatenate([a, b])))


# :234-247
def test_transitions_to_and_from_unity_use_the_same_ramp():
    p = oracle.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(preampDB=6))
    run(p, 960, 1, 1)
    p.setTarget(None)
    l, r = run(p, 960, 1, 1)
    gain = np.float32(10.0 ** (6.0 / 20.0))
    assert abs(l[0] - (gain - (gain - 1) / np.float32(960))) <= 1e-5
    assert abs(l[-1] - 1) <= 1e-5 and abs(r[-1] - 1) <= 1e-5


# :249-265
def test_rapid_publication_queues_newest_target():
    p = oracle.ParametricEqualizerProcessor(48000)
    pos, neg = np.float32(10 ** (6 / 20)), np.float32(10 ** (-6 / 20))
    p.setTarget(dict(preampDB=6))
    run(p, 480, 1, 1)
    p.setTarget(dict(preampDB=-6))
    l, _ = run(p, 480, 1, 1)
    assert abs(l[-1] - pos) <= 1e-5
    l, r = run(p, 960, 1, 1)
    assert abs(l[-1] - neg) <= 1e-5 and np.all(np.isfinite(l)) and np.all(np.isfinite(r))


# :267-288
def test_retirement_pressure_defers_transition_until_control_drain():
    p = oracle.ParametricEqualizerProcessor(48000)
    g1, g2, g3 = (np.float32(10 ** (d / 20)) for d in (6, -6, 12))
    p.setTarget(dict(preampDB=6))
    run(p, 960, 1, 1)
    p.setTarget(dict(preampDB=-6))
    second, _ = run(p, 960, 1, 1)
    assert abs(second[-1] - g2) <= 1e-5
    p.setTarget(dict(preampDB=12))
    held, _ = run(p, 960, 1, 1)
    assert abs(held[-1] - g2) <= 1e-5
    p.drainRetiredStates()
    newest, _ = run(p, 960, 1, 1)
    assert abs(newest[-1] - g3) <= 1e-5
    assert abs(second[0] - (g1 + (g2 - g1) / np.float32(960))) <= 1e-5


# :285-302
def test_render_callback_keeps_prior_target_when_publication_lock_contended():
    p = oracle.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(preampDB=6))
    p.holdPublicationLock(True)
    l, r = run(p, 128, 1, 2)
    p.holdPublicationLock(False)
    assert np.all(l == 1) and np.all(r == 2)


# :304-315
def test_reset_clears_published_state_histories():
    p = oracle.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(filters=[make_filter("peaking", 1000, 6, 0.707)]))
    run(p, 960, 1, 1)
    p.reset()
    p.setTarget(None)
    run(p, 960, 1, 1)
    l, r = run(p, 1, 0, 0)
    assert l.tolist() == [0] and r.tolist() == [0]


# :359-394
def test_reference_fixture_matches_representative_curve(eq_fixture_bytes):
    k = KAT["fixture_curve_db"]
    definition = oracle.parse_equalizer_apo(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    assert sum(f["isEnabled"] for f in definition["filters"]) == 10
    fs, n, discard = k["sampleRate"], k["frameCount"], k["discardCount"]
    for f, want in k["points"]:
        state = oracle.ParametricEqualizerProcessor.prepare(definition, fs)
        x = np.sin(2 * np.pi * f * np.arange(n) / fs).astype(np.float32)
        l, r = state.process(x, x)
        rms_in = math.sqrt(float(np.mean(x[discard:].astype(np.float64) ** 2)))
        rms_out = math.sqrt(float(np.mean(l[discard:].astype(np.float64) ** 2)))
        assert np.all(np.isfinite(l)) and np.all(np.isfinite(r))
        assert abs(20 * math.log10(rms_out / rms_in) - want) <= k["tolerance_db"]
