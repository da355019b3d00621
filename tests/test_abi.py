"""CPU-only checks of the drop-in boundary: libairwave_cuda.so loads, exports every symbol
include/airwave_cuda.h declares, and its host-side (non-CUDA) entry points agree with the oracle
and with the reference's known-answer vectors.  No kernel is launched here."""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import airwave_b200 as aw
import oracle
from conftest import GOLDEN, ROOT

KAT = json.load(open(os.path.join(GOLDEN, "kat_reference.json")))


def test_library_is_built_and_exports_every_declared_symbol():
    symbols = aw.declared_symbols()
    assert len(symbols) >= 40 and len(set(symbols)) == len(symbols)
    L = aw.lib()
    for name in symbols:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", aw.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(symbols) <= exported
    # nothing but the C ABI leaks out of the shared object
    assert all(s.startswith("aw_") for s in exported), sorted(s for s in exported if not s.startswith("aw_"))[:5]


def test_header_is_plain_c():
    src = '#include "airwave_cuda.h"\nint main(void) { aw_engine_config c; (void)c; return AW_OK; }\n'
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src, text=True, check=True)


def test_no_torch_types_and_no_oracle_in_the_product():
    header = open(os.path.join(ROOT, "include", "airwave_cuda.h")).read()
    assert "torch" not in header.lower() and "at::" not in header
    for root, _, files in os.walk(os.path.join(ROOT, "airwave_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp", ".swift")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "airwave_oracle" not in text, f


@pytest.mark.parametrize("case", KAT["biquad_coefficients"]["cases"])
def test_biquad_golden_coefficients_through_the_abi(case):
    c = aw.BiquadCoefficientBuilder.make(case["type"], case["gainDB"], case["frequencyHz"], case["q"], case["sampleRate"])
    assert np.abs(c - np.asarray(case["expected"])).max() <= 1e-12
    assert np.array_equal(c, oracle.biquad_make(case["type"], case["gainDB"], case["frequencyHz"], case["q"], case["sampleRate"]))


def test_biquad_error_codes_match_the_reference_order():
    mk = aw.BiquadCoefficientBuilder.make
    for args, name in [(("peaking", 1, 1000, 1, 0), "invalidSampleRate"), (("peaking", float("nan"), 1000, 1, 48000), "nonFiniteInput"),
                       (("peaking", 1, 24000, 1, 48000), "invalidFrequency"), (("peaking", 1, 0, 1, 48000), "invalidFrequency"),
                       (("peaking", 1, 1000, 0, 48000), "invalidQ")]:
        with pytest.raises(aw.BiquadCoefficientError) as e:
            mk(*args)
        assert e.value.name == name
        with pytest.raises(oracle.BiquadCoefficientError) as eo:
            oracle.biquad_make(*args)
        assert eo.value.name == name


def test_parser_matches_reference_fixture_and_oracle(eq_fixture_bytes, golden_dir):
    k = KAT["parser_fixture"]
    d = aw.EqualizerAPOParser.parse(eq_fixture_bytes, "CCA CRA ParametricEq.txt")
    assert d["preampDB"] == k["preampDB"]
    assert [f["frequencyHz"] for f in d["filters"]] == k["frequencyHz"]
    assert [f["gainDB"] for f in d["filters"]] == k["gainDB"]
    assert [f["q"] for f in d["filters"]] == k["q"]
    assert d == oracle.parse_equalizer_apo(eq_fixture_bytes, "x")
    for name in ["Bass Booster", "Bass Reducer", "Treble Booster", "Treble Reducer", "Vocal Booster"]:
        data = open(os.path.join(golden_dir, "eq", name + ".txt"), "rb").read()
        assert aw.EqualizerAPOParser.parse(data, name) == oracle.parse_equalizer_apo(data, name)


PARSER_CASES = [
    "# comment\nPreamp: -2.5 dB\nFilter 7: ON PK Fc 1000 Hz Gain 3.25 dB Q 1.20\nFilter: off LSC Fc 80 Hz Gain -1 dB Q 0.7\nFilter 9: ON HSC Fc 10000 Hz Gain -2 dB Q 0.70",
    "﻿  pReAmP : 1e0 dB\r\n\t# ignored\r\n fIlTeR 1 : oN pK Fc 440 Hz gAiN 2 dB q 1\r\n",
    "Filter 1: OFF PK Fc 440 Hz Gain 2 dB Q 1",
    "Filter 1: ON PK Fc 440 Hz Gain 2 dB Q 1",
    "Preamp: 1 dB\nPreamp: 2 dB\nFilter 1: ON PK Fc 440 Hz Gain 2 dB\nInclude: other.txt",
    "Preamp: NaN dB\nFilter 1: ON PK Fc 0 Hz Gain inf dB Q -1",
    "\n".join(f"Filter {i}: ON PK Fc {i} Hz Gain 1 dB Q 1" for i in range(1, 66)),
    "Preamp 3 dB\nPreamp: 3dB\nFilter1: ON PK Fc 1 Hz Gain 1 dB Q 1\nFilter 2:ON PK Fc 1 Hz Gain 1 dB Q 1",
    "Preamp:3 dB\nFilter 12 : ON LSC Fc 1e2 Hz Gain -.5 dB Q 7.\nFilter 3: ON PK Fc 0x10 Hz Gain 1 dB Q 1",
]


@pytest.mark.parametrize("text", PARSER_CASES)
def test_parser_agrees_with_oracle_on_reference_test_inputs(text):
    data = text.encode("utf-8")
    try:
        want = oracle.parse_equalizer_apo(data, "t.txt")
    except oracle.EqualizerParseError as e:
        with pytest.raises(aw.EqualizerParseError) as got:
            aw.EqualizerAPOParser.parse(data, "t.txt")
        assert got.value.issues == e.issues
        assert str(got.value) == str(e)
        return
    assert aw.EqualizerAPOParser.parse(data, "t.txt") == want


def test_parser_rejects_oversized_and_invalid_utf8():
    with pytest.raises(aw.EqualizerParseError) as e:
        aw.EqualizerAPOParser.parse(b" " * (1_048_576 + 1), "large.txt")
    assert "1 MiB" in str(e.value) and e.value.filename == "large.txt"
    with pytest.raises(aw.EqualizerParseError) as e:
        aw.EqualizerAPOParser.parse(b"Preamp: 1 dB\xff", "bad.txt")
    assert "UTF-8" in str(e.value)


@pytest.mark.parametrize("name", ["NeutralSH1.0", "RoomSH1.0", "StageSH1.0"])
def test_wav_loader_matches_oracle_on_bundled_presets(name, hrtf_path):
    w = aw.WAVLoader.load(hrtf_path(name))
    o = oracle.load_wav(hrtf_path(name))
    assert (w.sampleRate, w.channelCount, w.frameCount) == (o.sampleRate, o.channelCount, o.frameCount) == (48000.0, 14, 4320)
    assert np.array_equal(w.audioData, o.audioData)


def _wav(tag, bits, channels, payload, extensible=False):
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, 44100, 44100 * block, block, bits)
    if extensible:
        fmt += struct.pack("<HHIH", 22, bits, 3, tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00"
    body += b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_wav_sample_formats_and_errors():
    cases = [(1, 16, 2, np.array([[-32768, 16384], [32767, 0]], "<i2").tobytes(), False),
             (1, 32, 1, np.array([[-2147483648], [1 << 30]], "<i4").tobytes(), False),
             (1, 24, 1, bytes([0, 0, 0x80, 0, 0, 0x40]), False),
             (3, 32, 2, np.array([[0.25, -0.5]], "<f4").tobytes(), True)]
    for tag, bits, ch, payload, ext in cases:
        data = _wav(tag, bits, ch, payload, ext)
        assert np.array_equal(aw.WAVLoader.load(data).audioData, oracle.load_wav(data).audioData)
    for bad, status in [(b"not a wav file at all", 30), (_wav(1, 16, 2, b""), 32), (_wav(1, 8, 1, b"\x01\x02"), 33)]:
        with pytest.raises(aw.AirwaveError) as e:
            aw.WAVLoader.load(bad)
        assert e.value.status == status

def test_wav_loader_rejects_hostile_headers():
    """The header is untrusted: block_align smaller than a frame's samples would make the loader read past the data chunk
    (the reference gets this check from AVAudioFile); such files are refused, never read out of bounds."""
    def hostile(tag, bits, channels, block_align, payload):
        fmt = struct.pack("<HHIIHH", tag, channels, 48000, 48000 * block_align, block_align, bits)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
        return b"RIFF" + struct.pack("<I", len(body)) + body
    for tag, bits, ch, ba in [(3, 32, 14, 1), (3, 64, 4096, 2), (1, 16, 2, 3), (1, 24, 65535, 1)]:
        with pytest.raises(aw.AirwaveError) as e:
            aw.WAVLoader.load(hostile(tag, bits, ch, ba, b"\x00" * 64))
        assert e.value.status == 33, (tag, bits, ch, ba)
    # padding inside a frame (block_align larger than the samples) is legal and decodes like the tight layout
    tight = np.array([[0.25, -0.5], [0.125, 1.0]], "<f4")
    padded = b"".join(row.tobytes() + b"\xff" * 4 for row in tight)
    w = aw.WAVLoader.load(hostile(3, 32, 2, 12, padded))
    assert np.array_equal(w.audioData, tight.T)



def test_layouts_and_hesuvi_maps():
    assert aw.InputLayout.surround71().channels == oracle.InputLayout.surround71.channels
    assert aw.InputLayout.stereo().channels == ["FL", "FR"]
    assert aw.InputLayout.atmos714().channels == oracle.InputLayout.atmos714.channels
    sp = aw.InputLayout.atmos714().channels
    for n, fn in [(14, "hesuvi14Channel"), (7, "hesuvi7Channel")]:
        got = getattr(aw.HRIRChannelMap, fn)(sp).mapping
        want = getattr(oracle.HRIRChannelMap, fn)(sp).mapping
        assert got == want and len(got) == 8
    for key in ("hesuvi14_map", "hesuvi7_map"):
        m = (aw.HRIRChannelMap.hesuvi14Channel if "14" in key else aw.HRIRChannelMap.hesuvi7Channel)(sp[:8])
        for s in sp[:8]:
            assert list(m.getIndices(s)) == KAT[key][s]
    text = "# c\nFL = 0, 1\nR=8,7\n; x\nbogus\nSUB = 6, 13\nXX = 1, 2\nSL = 2\nrl = 4 , 5\nTBR=3,x\n"
    got = aw.HRIRChannelMap.parseHeSuViFormat(text).mapping
    want = {k: v for k, v in oracle.HRIRChannelMap.parseHeSuViFormat(text).mapping.items() if not k.startswith("custom:")}
    assert got == want == {"FL": (0, 1), "FR": (8, 7), "LFE": (6, 13), "BL": (4, 5)}


def test_resample_output_count():
    L = aw.lib()
    assert L.aw_resample_output_count(4320, 44100.0, 48000.0) == 4702 == oracle.resample_output_count(4320, 44100.0, 48000.0)
    assert L.aw_resample_output_count(4320, 48000.0, 48000.0) == 4320
    assert L.aw_resample_output_count(4320, 96000.0, 48000.0) == 2160


def test_create_without_cuda_fails_loudly_instead_of_falling_back():
    if aw.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(aw.AirwaveError) as e:
        aw.BinauralEngine(1, 2, 512)
    assert e.value.status == 2   # AW_ERR_CUDA
    with pytest.raises(aw.AirwaveError):
        aw.HRIRBank(np.ones((2, 8), np.float32), 48000.0, 48000.0, [0], [1], 8)


def test_cpp_mirror_header_compiles_and_links_against_the_abi(tmp_path):
    """include/airwave.hpp (C++ mirror of ConvolutionEngine / RealtimeAudioProcessor) type-checks against the C ABI."""
    exe = tmp_path / "hpp_check"
    lib_dir = os.path.dirname(aw.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-o", str(exe), os.path.join(ROOT, "tests", "cpu", "hpp_compile_check.cpp"),
                    "-L", lib_dir, "-lairwave_cuda", f"-Wl,-rpath,{lib_dir}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_swift_shim_and_module_map_are_shipped():
    mm = open(os.path.join(ROOT, "include", "module.modulemap")).read()
    assert "module CAirwaveCUDA" in mm and "airwave_cuda.h" in mm
    swift = open(os.path.join(ROOT, "airwave_b200", "swift", "AirwaveCUDA.swift")).read()
    for name in ("class ConvolutionEngine", "class RealtimeAudioProcessor", "protocol StereoAudioProcessing", "import CAirwaveCUDA"):
        assert name in swift
    used = set(__import__("re").findall(r"\b(aw_[a-z0-9_]+)\(", swift)) - {"aw_engine_config"}   # struct initialiser, not a call
    assert used and used <= set(aw.declared_symbols())
    # the app-target file: the two effects AudioEffectGraph composes (AudioEffectGraph.swift:47-54), every C call declared in the header
    effects = open(os.path.join(ROOT, "airwave_b200", "swift", "AirwaveCUDAEffects.swift")).read()
    for name in ("class CUDASpatialEffect: AudioSpatialEffect", "class CUDAEqualizerEffect: AudioEqualizerEffect",
                 "func prepare(definition: EqualizerDefinition?, sampleRate: Double) throws",
                 "func setTarget(definition: EqualizerDefinition?) throws", "var isReady"):
        assert name in effects, name
    used = set(__import__("re").findall(r"\b(aw_[a-z0-9_]+)\(", effects)) - {"aw_engine_config", "aw_eq_filter"}
    assert used and used <= set(aw.declared_symbols()), used - set(aw.declared_symbols())


def test_realtime_path_gate():
    """The reference gates its render callback with a grep (scripts/check-audio-safety-invariants.sh:23-41: no Array growth,
    DispatchQueue, locks, print, Logger between BEGIN/END REALTIME CALLBACK).  Same gate for aw_engine_process*: between the
    BEGIN/END REALTIME PATH markers of aw_api.cu nothing may allocate, grow a container, lock, log or synchronise the device.
    Lines that only build an error message (set_error) are exempt: they run after the call has already failed."""
    import re
    text = open(os.path.join(ROOT, "airwave_b200", "csrc", "aw_api.cu")).read()
    regions = re.findall(r"// BEGIN REALTIME PATH(.*?)// END REALTIME PATH", text, re.S)
    assert len(regions) == 2, "markers missing"
    body = "\n".join(regions)
    for name in ("process_device_body", "process_block", "eq_process_machine", "aw_engine_process_stereo", "aw_engine_submit_device",
                 "aw_engine_wait", "eq_begin_call"):
        assert name + "(" in body, f"{name} is not inside the gated region"
    forbidden = [r"\bcudaMalloc", r"\bcudaFree\b", r"\bcudaHostAlloc", r"\bcudaFreeHost", r"\bnew\s", r"\bdelete\s", r"push_back", r"emplace_back",
                 r"\.resize\(", r"\.reserve\(", r"\.insert\(", r"\.erase\(", r"std::vector<", r"std::map", r"std::mutex", r"lock_guard",
                 r"\bprintf", r"std::cout", r"std::cerr", r"cudaDeviceSynchronize", r"cudaStreamCreate", r"cudaEventCreate\b",
                 r"std::to_string", r"\bmalloc\(", r"getenv"]
    bad = []
    for n, line in enumerate(body.splitlines()):
        code = line.split("//")[0]
        if "set_error(" in code:
            continue
        for pat in forbidden:
            if re.search(pat, code):
                bad.append((pat, line.strip()))
    assert not bad, bad
