"""GPU parity tests of the equalizer path through the C ABI: ports of
AirwaveTests/ParametricEqualizerProcessorTests.swift and AirwaveTests/AudioEffectGraphTests.swift run against
the CUDA path, plus CUDA vs the CPU oracle.  The EQ arithmetic is float64 with the reference's operation
order and no FMA contraction, so agreement with the oracle is required to be BIT-EXACT."""
import json
import math
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))
SEED = 0x41495257


@pytest.fixture(scope="module")
def aw():
    import airwave_b200
    assert airwave_b200.device_count() > 0
    return airwave_b200


def make_filter(type_, frequency, gain, q, enabled=True, line=1):
    return dict(type=type_, frequencyHz=frequency, gainDB=gain, q=q, isEnabled=enabled, sourceLine=line)


def run(p, n, lv, rv):
    return p.process(np.full(n, lv, np.float32), np.full(n, rv, np.float32))


# :87-107
def test_unity_and_preamp_only_states(aw):
    unity = aw.ParametricEqualizerProcessor.prepare(None, 48000)
    preamp = aw.ParametricEqualizerProcessor.prepare(dict(preampDB=6, filters=[]), 48000)
    l = np.array([0.25, -0.5, 1], np.float32)
    r = np.array([-0.75, 0.5, 0.125], np.float32)
    ul, ur = unity.process(l, r)
    pl, pr = preamp.process(l, r)
    gain = np.float32(10.0 ** (6.0 / 20.0))
    assert np.array_equal(ul, l) and np.array_equal(ur, r)
    for i in (0, 2):
        assert abs(pl[i] - l[i] * gain) <= 1e-6 and abs(pr[i] - r[i] * gain) <= 1e-6


# :109-133
def test_known_impulse_response_preserves_cascade_order(aw):
    k = KAT["cascade_impulse_response"]
    filters = [make_filter(f["type"], f["frequencyHz"], f["gainDB"], f["q"]) for f in k["filters"]]
    state = aw.ParametricEqualizerProcessor.prepare(dict(filters=filters), k["sampleRate"])
    l, r = state.process(k["left_input"], [0] * 6)
    assert np.abs(l - np.asarray(k["expected_left"], np.float32)).max() <= k["tolerance"]
    assert np.all(r == 0)


# :135-152
def test_disabled_filters_excluded_and_subnormal_flushed(aw):
    state = aw.ParametricEqualizerProcessor.prepare(dict(filters=[make_filter("peaking", 1000, 12, 0.7, enabled=False)]), 48000)
    l, r = state.process([1, 0], [1, 0])
    assert l.tolist() == [1, 0] and r.tolist() == [1, 0]
    active = aw.ParametricEqualizerProcessor.prepare(dict(filters=[make_filter("peaking", 1000, 6, 0.707)]), 48000)
    l, _ = active.process(np.array([np.float32(1e-45), 0], np.float32), [0, 0])
    assert l[0] != 0 and l[1] == 0


# :191-209
def test_preparation_rejects_invalid_inputs(aw):
    P = aw.ParametricEqualizerProcessor
    with pytest.raises(aw.ParametricEqualizerPreparationError) as e:
        P.prepare(None, 0)
    assert e.value.name == "invalidSampleRate"
    with pytest.raises(aw.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 24000, 1, 1)]), 48000)
    assert e.value.name == "invalidFilter" and e.value.index == 0 and e.value.filter_error == 2
    with pytest.raises(aw.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 1000, 1, 0)]), 48000)
    assert e.value.filter_error == 3
    with pytest.raises(aw.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(filters=[make_filter("peaking", 500 + i, 1, 1) for i in range(65)]), 48000)
    assert e.value.name == "tooManyFilters"
    with pytest.raises(aw.ParametricEqualizerPreparationError) as e:
        P.prepare(dict(preampDB=float("inf"), filters=[]), 48000)
    assert e.value.name == "nonFinitePreamp"


# :211-232
@pytest.mark.parametrize("fs", [44100.0, 48000.0, 96000.0])
def test_crossfade_uses_exact_twenty_millisecond_ramp(aw, fs):
    p = aw.ParametricEqualizerProcessor(fs, 4096)
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
def test_transitions_to_and_from_unity_use_the_same_ramp(aw):
    p = aw.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(preampDB=6))
    run(p, 960, 1, 1)
    p.setTarget(None)
    l, r = run(p, 960, 1, 1)
    gain = np.float32(10.0 ** (6.0 / 20.0))
    assert abs(l[0] - (gain - (gain - 1) / np.float32(960))) <= 1e-5
    assert abs(l[-1] - 1) <= 1e-5 and abs(r[-1] - 1) <= 1e-5


# :249-265
def test_rapid_publication_queues_newest_target(aw):
    p = aw.ParametricEqualizerProcessor(48000)
    pos, neg = np.float32(10 ** (6 / 20)), np.float32(10 ** (-6 / 20))
    p.setTarget(dict(preampDB=6))
    run(p, 480, 1, 1)
    p.setTarget(dict(preampDB=-6))
    l, _ = run(p, 480, 1, 1)
    assert abs(l[-1] - pos) <= 1e-5
    l, r = run(p, 960, 1, 1)
    assert abs(l[-1] - neg) <= 1e-5 and np.all(np.isfinite(l)) and np.all(np.isfinite(r))


# :267-288
def test_retirement_pressure_defers_transition_until_control_drain(aw):
    p = aw.ParametricEqualizerProcessor(48000)
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
def test_render_callback_keeps_prior_target_when_publication_lock_contended(aw):
    p = aw.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(preampDB=6))
    p.holdPublicationLock(True)
    l, r = run(p, 128, 1, 2)
    p.holdPublicationLock(False)
    assert np.all(l == 1) and np.all(r == 2)


# :304-315
def test_reset_clears_published_state_histories(aw):
    p = aw.ParametricEqualizerProcessor(48000)
    p.setTarget(dict(filters=[make_filter("peaking", 1000, 6, 0.707)]))
    run(p, 960, 1, 1)
    p.reset()
    p.setTarget(None)
    run(p, 960, 1, 1)
    l, r = run(p, 1, 0, 0)
    assert l.tolist() == [0] and r.tolist() == [0]


# :359-394
def test_reference_fixture_matches_representative_curve(aw, eq_fixture_bytes):
    k = KAT["fixture_curve_db"]
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    assert sum(f["isEnabled"] for f in definition["filters"]) == 10
    fs, n, discard = k["sampleRate"], k["frameCount"], k["discardCount"]
    for f, want in k["points"]:
        state = aw.ParametricEqualizerProcessor.prepare(definition, fs)
        x = np.sin(2 * np.pi * f * np.arange(n) / fs).astype(np.float32)
        l, r = state.process(x, x)
        rms_in = math.sqrt(float(np.mean(x[discard:].astype(np.float64) ** 2)))
        rms_out = math.sqrt(float(np.mean(l[discard:].astype(np.float64) ** 2)))
        assert np.all(np.isfinite(l)) and np.array_equal(l, r)
        assert abs(20 * math.log10(rms_out / rms_in) - want) <= k["tolerance_db"]
        ol, _ = oracle.ParametricEqualizerProcessor.prepare(definition, fs).process(x, x)
        assert np.array_equal(l, ol)   # bit-exact with the CPU restatement


def test_processor_is_bit_exact_with_oracle_across_transitions_and_callback_sizes(aw, eq_fixture_bytes):
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "f.txt")
    other = dict(preampDB=-3.0, filters=[make_filter("peaking" if i % 2 == 0 else "highShelf", 250 + i * 1000, (i % 3) - 1, 0.8)
                                         for i in range(10)])   # the reference's 10-filter measure{} workload (:317-357)
    g = aw.ParametricEqualizerProcessor(48000)
    o = oracle.ParametricEqualizerProcessor(48000)
    pos = 0
    script = {0: definition, 3: other, 4: None, 7: definition, 8: other}
    for i, size in enumerate([128, 512, 1024, 777, 1, 4096, 333, 960, 100, 2000, 4096]):
        if i in script:
            g.setTarget(script[i]); o.setTarget(script[i])
            g.drainRetiredStates(); o.drainRetiredStates()
        x = oracle.synth_block(SEED, [1], 2, pos, size)[0] * 3
        pos += size
        gl, gr = g.process(x[0], x[1])
        ol, orr = o.process(x[0], x[1])
        assert np.array_equal(gl, ol) and np.array_equal(gr, orr), (i, size)
        if i == 5:
            g.reset(); o.reset()


def test_batched_eq_per_range_targets_match_single_streams(aw, eq_fixture_bytes):
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "f.txt")
    bass = aw.EqualizerAPOParser.parse(open(os.path.join(GOLDEN, "eq", "Bass Booster.txt"), "rb").read(), "b")
    n = 70
    eng = aw.BinauralEngine(n, 2, 256, 48000.0, 1024)
    eng.eq_prepare(definition, 0, 40)
    eng.eq_prepare(bass, 40, 20)          # streams 60..69: EQ never prepared -> bypass
    x = oracle.synth_block(SEED, range(n), 2, 0, 2048)
    y = np.concatenate([eng.process(x[:, :, :1000]), eng.process(x[:, :, 1000:1024]), eng.process(x[:, :, 1024:])], axis=2)
    assert np.array_equal(y[60:], x[60:])
    for idx, d in [(0, definition), (39, definition), (40, bass), (59, bass)]:
        o = oracle.ParametricEqualizerProcessor(48000)
        o.setTarget(d)
        ol, orr = [], []
        for a, b in [(0, 1000), (1000, 1024), (1024, 2048)]:
            l, r = o.process(x[idx, 0, a:b], x[idx, 1, a:b])
            ol.append(l); orr.append(r)
        assert np.array_equal(y[idx, 0], np.concatenate(ol)) and np.array_equal(y[idx, 1], np.concatenate(orr))


# ---- AudioEffectGraphTests.swift (production effects) ------------------------------------------------
def test_graph_neither_effect_copies_stereo_and_duplicates_mono(aw):   # :5-22
    g = aw.AudioEffectGraph(8)
    res = g.prepare(48000, None)
    assert res.noEffectCanRun
    l, r = g.process([1, 2], [3, 4])
    assert l.tolist() == [1, 2] and r.tolist() == [3, 4]
    l, r = g.process([5, 6], None)
    assert l.tolist() == [5, 6] and r.tolist() == [5, 6]


def test_graph_production_equalizer_uses_output_rate_and_rejects_nyquist(aw):   # :96-111
    g = aw.AudioEffectGraph(8)
    invalid = dict(filters=[make_filter("peaking", 23000, 1, 1, line=31)])
    res = g.prepare(44100, invalid)
    assert res.noEffectCanRun and res.equalizerWarning["filterLine"] == 31 and "Nyquist" in res.equalizerWarning["reason"]
    res = g.prepare(96000, dict(preampDB=3, filters=[make_filter("peaking", 23000, 1, 1, line=31)]))
    assert res.runnableEffects == {"equalizer"} and res.equalizerWarning is None


def test_graph_equalizer_can_reenable_after_none_selection(aw):   # :113-163 (shape of the scenario)
    g = aw.AudioEffectGraph(4096)
    a = dict(preampDB=6)
    gain = np.float32(10 ** (6 / 20))
    assert g.prepare(48000, a).runnableEffects == {"equalizer"}
    l, _ = g.process(np.ones(960, np.float32), np.ones(960, np.float32))
    assert abs(l[-1] - gain) <= 1e-5
    assert g.updateEqualizer(None).noEffectCanRun            # unity ramp stays in the callback path
    l, _ = g.process(np.ones(960, np.float32), np.ones(960, np.float32))
    assert abs(l[0] - (gain - (gain - 1) / np.float32(960))) <= 1e-5 and abs(l[-1] - 1) <= 1e-5
    assert g.updateEqualizer(a).runnableEffects == {"equalizer"}
    l, _ = g.process(np.ones(960, np.float32), np.ones(960, np.float32))
    assert abs(l[-1] - gain) <= 1e-5
    assert g.prepare(48000, None).noEffectCanRun             # prepare(nil) bypasses the processor for a new pipeline
    l, r = g.process(np.ones(8, np.float32), None)
    assert np.all(l == 1) and np.all(r == 1)


def test_graph_update_before_prepare_reports_unavailable(aw):
    g = aw.AudioEffectGraph(64)
    res = g.updateEqualizer(dict(preampDB=1))
    assert "not been prepared" in res.equalizerWarning["reason"]


def test_graph_spatial_then_equalizer_matches_oracle_chain(aw, hrtf_path, eq_fixture_bytes):   # order of :54-70
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "f.txt")
    g = aw.AudioEffectGraph(4096, 512)
    g.activatePreset(aw.WAVLoader.load(hrtf_path("NeutralSH1.0")), 48000.0)
    assert g.prepare(48000.0, definition).runnableEffects == {"spatial", "equalizer"}
    wav_o = oracle.load_wav(hrtf_path("NeutralSH1.0"))
    spatial = oracle.RealtimeAudioProcessor(oracle.activate_preset(wav_o, 48000.0, oracle.InputLayout.stereo, 512), 512, 4096)
    spatial.isReady = True
    eq = oracle.ParametricEqualizerProcessor(48000.0)
    eq.setTarget(definition)
    model = oracle.AudioEffectGraphModel(spatial, eq, 4096)
    model.equalizerActive = True
    pos = 0
    for size in [512, 300, 1024, 4096, 212, 512]:
        x = oracle.synth_block(SEED, [4], 2, pos, size)[0]
        pos += size
        gl, gr = g.process(x[0], x[1])
        ol, orr = model.process(x[0], x[1])
        assert np.abs(gl - ol).max() <= 1e-5 and np.abs(gr - orr).max() <= 1e-5
    g.deactivatePreset()
    assert g.prepare(48000.0, None).noEffectCanRun


def test_overlapped_equalizer_changes_no_bit(aw, hrtf_path, eq_fixture_bytes):
    """AW_ENGINE_OVERLAP_EQ: the equalizer of call j runs on an internal stream next to the convolution of call j+1
    (AudioEffectGraph's order spatial -> equalizer, AudioEffectGraph.swift:195-210, is kept per call).  Same samples as the
    serial engine through every entry point: device calls (two alternating output buffers, results picked up one call later as
    the contract says, the last after aw_engine_flush), the synchronous host call, and the pipelined submit/wait path — with a
    live target change in the middle, so the crossfade's two voices and the reset of the new voice run on the internal stream."""
    import torch
    n, S, B, calls, frames = 300, 8, 256, 7, 1024
    definition = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    other = dict(preampDB=-3.0, filters=[make_filter("peaking", 1000.0, 4.0, 1.2), make_filter("lowShelf", 120.0, -2.0, 0.7)])
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), 48000.0, aw.InputLayout.surround71(), B)
    x = oracle.synth_block(SEED, [7 + 3 * i for i in range(n)], S, 0, calls * frames)
    xd = torch.from_numpy(x).cuda()

    def device_run(overlap):
        eng = aw.BinauralEngine(n, S, B, 48000.0, max_frames_per_call=frames, max_partitions=bank.partitions, overlap_eq=overlap)
        eng.set_bank(bank)
        eng.eq_prepare(definition)
        stream = torch.cuda.ExternalStream(eng.cuda_stream)
        ybuf = [torch.empty((n, 2, frames), dtype=torch.float32, device="cuda") for _ in range(2)]
        outs = [None] * calls
        for j in range(calls):
            if j == 3:
                eng.eq_update(other)
            xin = xd[:, :, j * frames:(j + 1) * frames].contiguous()
            torch.cuda.synchronize()
            eng.process_device(xin.data_ptr(), S * frames, frames, ybuf[j % 2].data_ptr(), 2 * frames, frames, frames)
            if overlap:
                if j > 0:                                  # call j-1 is complete in stream order once call j is
                    with torch.cuda.stream(stream):
                        outs[j - 1] = ybuf[(j - 1) % 2].clone()
            else:
                with torch.cuda.stream(stream):
                    outs[j] = ybuf[j % 2].clone()
        if overlap:
            eng.flush()
            with torch.cuda.stream(stream):
                outs[calls - 1] = ybuf[(calls - 1) % 2].clone()
        torch.cuda.synchronize()
        eng.close()
        return np.concatenate([o.cpu().numpy() for o in outs], axis=2)

    serial = device_run(False)
    assert np.array_equal(device_run(True), serial)

    def host_run(overlap, pipelined):
        eng = aw.BinauralEngine(n, S, B, 48000.0, max_frames_per_call=frames, max_partitions=bank.partitions, overlap_eq=overlap,
                                pipelined=pipelined)
        eng.set_bank(bank)
        eng.eq_prepare(definition)
        if not pipelined:
            outs = []
            for j in range(calls):
                if j == 3:
                    eng.eq_update(other)
                outs.append(eng.process(np.ascontiguousarray(x[:, :, j * frames:(j + 1) * frames])))
            eng.close()
            return np.concatenate(outs, axis=2)
        hin = [aw.PinnedBuffer((n, S, frames)) for _ in range(calls)]
        hout = [aw.PinnedBuffer((n, 2, frames)) for _ in range(calls)]
        for j in range(calls):
            hin[j].array[...] = x[:, :, j * frames:(j + 1) * frames]
        for j in range(calls):
            if j == 3:
                eng.eq_update(other)
            eng.submit(hin[j].array.ctypes.data, hout[j].array.ctypes.data, frames)
        eng.wait()
        y = np.concatenate([h.array.copy() for h in hout], axis=2)
        eng.close()
        return y

    assert np.array_equal(host_run(True, False), serial)
    assert np.array_equal(host_run(True, True), serial)
    assert np.array_equal(host_run(False, True), serial)


def test_equalizer_state_pool_grows_with_the_number_of_ranges(aw, hrtf_path):
    """Per-device profiles: every stream range holds its own ParametricEqualizerState objects (DeviceProfileManager.swift:4-12).
    1,500 single-stream ranges with their own equalizers need more than the 1,024 program slots an engine starts with: the pool
    grows in the control path, and what each stream renders is what a one-range engine renders for the same definition."""
    n, B = 1500, 256
    bank = aw.HRIRBank.from_wav(aw.WAVLoader.load(hrtf_path("RoomSH1.0")), 48000.0, aw.InputLayout.stereo(), B)
    eng = aw.BinauralEngine(n, 2, B, 48000.0, max_frames_per_call=1024, max_partitions=bank.partitions)
    eng.set_bank(bank)
    defs = [dict(preampDB=-1.0 - (i % 7), filters=[make_filter("peaking", 200.0 + 13.0 * (i % 50), 3.0, 1.0)]) for i in range(n)]
    for i in range(n):
        eng.eq_install_state(defs[i], i, 1)
    x = oracle.synth_block(SEED, [3] * n, 2, 0, 2048)
    y = np.concatenate([eng.process(np.ascontiguousarray(x[:, :, a:a + 1024])) for a in (0, 1024)], axis=2)
    eng.close()
    for i in (0, 6, 511, 1023, 1024, 1499):
        ref_eng = aw.BinauralEngine(1, 2, B, 48000.0, max_frames_per_call=1024, max_partitions=bank.partitions)
        ref_eng.set_bank(bank)
        ref_eng.eq_install_state(defs[i])
        ref = np.concatenate([ref_eng.process(np.ascontiguousarray(x[:1, :, a:a + 1024])) for a in (0, 1024)], axis=2)
        ref_eng.close()
        assert np.array_equal(y[i], ref[0]), i
