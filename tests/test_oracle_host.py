"""Pins the oracle's host-side restatement: EqualizerAPOParser (AirwaveTests/EqualizerAPOParserTests.swift),
HeSuVi channel maps (Airwave/VirtualSpeaker.swift:224-297), WAV loading of the bundled presets and
AudioEffectGraph routing (AirwaveTests/AudioEffectGraphTests.swift:5-70)."""
import json
import os
import struct

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))


def parse(text, name="test.txt"):
    return oracle.parse_equalizer_apo(text.encode("utf-8") if isinstance(text, str) else text, name)


def parse_error(text, name="test.txt"):
    with pytest.raises(oracle.EqualizerParseError) as e:
        parse(text, name)
    return e.value


# EqualizerAPOParserTests.swift:7-27
def test_reference_fixture_parses_exactly(eq_fixture_bytes):
    k = KAT["parser_fixture"]
    d = parse(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    assert d["preampDB"] == k["preampDB"] and len(d["filters"]) == 10
    assert d["filters"][0]["type"] == "lowShelf" and d["filters"][-1]["type"] == "highShelf"
    assert all(f["isEnabled"] for f in d["filters"])
    assert [f["frequencyHz"] for f in d["filters"]] == k["frequencyHz"]
    assert [f["gainDB"] for f in d["filters"]] == k["gainDB"]
    assert [f["q"] for f in d["filters"]] == k["q"]


# :29-44
def test_parses_supported_definition_in_source_order():
    d = parse("# comment\nPreamp: -2.5 dB\nFilter 7: ON PK Fc 1000 Hz Gain 3.25 dB Q 1.20\n"
              "Filter: off LSC Fc 80 Hz Gain -1 dB Q 0.7\nFilter 9: ON HSC Fc 10000 Hz Gain -2 dB Q 0.70")
    assert d["preampDB"] == -2.5
    assert [f["sourceLine"] for f in d["filters"]] == [3, 4, 5]
    assert [f["sourceNumber"] for f in d["filters"]] == [7, None, 9]
    assert [f["isEnabled"] for f in d["filters"]] == [True, False, True]
    assert [f["type"] for f in d["filters"]] == ["peaking", "lowShelf", "highShelf"]
    assert [f["frequencyHz"] for f in d["filters"]] == [1000, 80, 10000]


# :46-54
def test_accepts_bom_crlf_whitespace_case_and_comments():
    d = parse("﻿  pReAmP : 1e0 dB\r\n\t# ignored\r\n fIlTeR 1 : oN pK Fc 440 Hz gAiN 2 dB q 1\r\n", "mixed.txt")
    assert d["preampDB"] == 1 and len(d["filters"]) == 1 and d["filters"][0]["gainDB"] == 2


# :56-67
def test_off_filter_does_not_make_configuration_effective():
    err = parse_error("Filter 1: OFF PK Fc 440 Hz Gain 2 dB Q 1")
    assert any("effective" in r for _, r in err.issues)
    d = parse("Filter 1: ON PK Fc 440 Hz Gain 2 dB Q 1")
    assert d["preampDB"] == 0 and d["filters"][0]["isEnabled"]


# :69-83
def test_rejects_malformed_unsupported_and_duplicate_directives():
    err = parse_error("Preamp: 1 dB\nPreamp: 2 dB\nFilter 1: ON PK Fc 440 Hz Gain 2 dB\nInclude: other.txt", "bad.txt")
    assert err.filename == "bad.txt"
    assert any(ln == 2 and "duplicate" in r for ln, r in err.issues)
    assert any(ln == 3 and "malformed" in r for ln, r in err.issues)
    assert any(ln == 4 and "unsupported" in r for ln, r in err.issues)


# :85-98
def test_rejects_non_finite_non_positive_and_too_many_filters():
    err = parse_error("Preamp: NaN dB\nFilter 1: ON PK Fc 0 Hz Gain inf dB Q -1")
    reasons = [r for _, r in err.issues]
    assert any("finite" in r for r in reasons) and any("frequency" in r for r in reasons) and any("Q" in r for r in reasons)
    many = "\n".join(f"Filter {i}: ON PK Fc {i} Hz Gain 1 dB Q 1" for i in range(1, 66))
    assert any("64" in r for _, r in parse_error(many).issues)


# :100-105
def test_rejects_oversized_data():
    err = parse_error(b" " * (1_048_576 + 1), "large.txt")
    assert err.filename == "large.txt" and any("1 MiB" in r for _, r in err.issues)


def test_bundled_eq_presets_parse(golden_dir):
    for name in ["Bass Booster", "Bass Reducer", "Treble Booster", "Treble Reducer", "Vocal Booster"]:
        d = parse(open(os.path.join(golden_dir, "eq", name + ".txt"), "rb").read(), name)
        assert d["filters"] and all(f["isEnabled"] for f in d["filters"])


# VirtualSpeaker.swift:224-297
def test_hesuvi_channel_maps():
    speakers = oracle.InputLayout.surround71.channels
    assert speakers == ["FL", "FR", "FC", "LFE", "BL", "BR", "SL", "SR"]
    m14 = oracle.HRIRChannelMap.hesuvi14Channel(speakers)
    m7 = oracle.HRIRChannelMap.hesuvi7Channel(speakers)
    for sp in speakers:
        assert list(m14.getIndices(sp)) == KAT["hesuvi14_map"][sp]
        assert list(m7.getIndices(sp)) == KAT["hesuvi7_map"][sp]
    assert m14.getIndices("TFL") is None
    parsed = oracle.HRIRChannelMap.parseHeSuViFormat("# c\nFL = 0, 1\nR=8,7\n; x\nbogus\nSUB = 6, 13\nXX = 1, 2\nSL = 2\n")
    assert parsed.mapping == {"FL": (0, 1), "FR": (8, 7), "LFE": (6, 13), "custom:XX": (1, 2)}
    assert oracle.InputLayout.detect(8).name == "7.1 Surround" and len(oracle.InputLayout.detect(5).channels) == 5


@pytest.mark.parametrize("name", ["NeutralSH1.0", "RoomSH1.0", "StageSH1.0"])
def test_bundled_hrir_presets_load(name, hrtf_path):
    w = oracle.load_wav(hrtf_path(name))
    assert (w.sampleRate, w.channelCount, w.frameCount) == (48000.0, 14, 4320)
    assert w.audioData.shape == (14, 4320) and np.all(np.abs(w.audioData).max(axis=1) > 0)
    e = (w.audioData.astype(np.float64) ** 2).sum(axis=1)
    # ipsilateral channels (0, 7) carry more energy than contralateral ones (1, 8): second witness of the order
    assert e[0] > e[1] and e[7] > e[8]


def _wav(tag, bits, channels, frames_bytes, extensible=False):
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, 44100, 44100 * block, block, bits)
    if extensible:
        fmt += struct.pack("<HHIH", 22, bits, 3, tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00"
    body += b"data" + struct.pack("<I", len(frames_bytes)) + frames_bytes
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_wav_sample_formats():  # WAVLoader.swift:66-91
    i16 = np.array([[-32768, 16384], [32767, 0]], "<i2")
    w = oracle.load_wav(_wav(1, 16, 2, i16.tobytes()))
    assert w.channelCount == 2 and w.frameCount == 2 and w.sampleRate == 44100.0
    assert w.audioData.tolist() == [[-1.0, 32767 / 32768], [0.5, 0.0]]
    i32 = np.array([[-2147483648], [1 << 30]], "<i4")
    assert oracle.load_wav(_wav(1, 32, 1, i32.tobytes())).audioData.tolist() == [[-1.0, 0.5]]
    f32 = np.array([[0.25, -0.5]], "<f4")
    assert oracle.load_wav(_wav(3, 32, 2, f32.tobytes(), extensible=True)).audioData.tolist() == [[0.25], [-0.5]]
    with pytest.raises(oracle.WAVError):
        oracle.load_wav(_wav(1, 16, 2, b""))
    with pytest.raises(oracle.WAVError):
        oracle.load_wav(b"not a wav file at all")


# AudioEffectGraphTests.swift:5-70
class SpatialSpy:
    def __init__(self, isReady, offset=0.0):
        self.isReady, self.offset, self.count = isReady, offset, 0

    def process(self, l, r):
        self.count += 1
        return l + self.offset, (l if r is None else r) + self.offset


class EqSpy:
    def __init__(self, multiplier=1.0):
        self.multiplier, self.count = multiplier, 0

    def process(self, l, r):
        self.count += 1
        return l * self.multiplier, (l if r is None else r) * self.multiplier


def test_graph_routing_modes():
    g = oracle.AudioEffectGraphModel(SpatialSpy(False), EqSpy(), 8)
    l, r = g.process([1, 2], [3, 4])
    assert l.tolist() == [1, 2] and r.tolist() == [3, 4]
    l, r = g.process([5, 6], None)
    assert l.tolist() == [5, 6] and r.tolist() == [5, 6]
    sp, eq = SpatialSpy(True, 10), EqSpy()
    g = oracle.AudioEffectGraphModel(sp, eq, 8)
    l, r = g.process([1], [2])
    assert (l.tolist(), r.tolist(), sp.count, eq.count) == ([11], [12], 1, 0)
    eq = EqSpy(2)
    g = oracle.AudioEffectGraphModel(SpatialSpy(False), eq, 8)
    g.equalizerActive = True
    l, r = g.process([1], None)
    assert (l.tolist(), r.tolist(), eq.count) == ([2], [2], 1)
    sp, eq = SpatialSpy(True, 10), EqSpy(2)
    g = oracle.AudioEffectGraphModel(sp, eq, 8)
    g.equalizerActive = True
    l, r = g.process([1], [2])
    assert (l.tolist(), r.tolist()) == ([22], [24])  # (1+10)*2: spatial then EQ
