"""Pure-Python restatement of the reference's setup-time host logic (TEST INFRASTRUCTURE ONLY).

WAVLoader, InputLayout, HRIRChannelMap, EqualizerAPOParser, HRIRManager activation build loop and
AudioEffectGraph routing.  Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
import re
import struct

import numpy as np

from . import binding as _b

__all__ = [
    "WAVData", "WAVError", "load_wav", "InputLayout", "HRIRChannelMap", "HRIRError",
    "EqualizerParseError", "parse_equalizer_apo", "activate_preset", "hrir_matrix",
    "AudioEffectGraphModel",
]


# ------------------------------------------------------------------------------------------------
# WAVLoader.load  (WAVLoader.swift:26-99).  AVAudioFile is replaced by a plain RIFF/WAVE parser
# (SURVEY.md Q13): PCM int16 -> /32768 (:78), int32 -> /2147483648 (:86), float32 as is (:66-71);
# int24 and WAVE_FORMAT_EXTENSIBLE are what AVAudioFile would decode to float.
# ------------------------------------------------------------------------------------------------
class WAVError(Exception):
    pass


class WAVData:
    def __init__(self, sampleRate: float, channelCount: int, frameCount: int, audioData: np.ndarray):
        self.sampleRate = sampleRate
        self.channelCount = channelCount
        self.frameCount = frameCount
        self.audioData = audioData  # [channel][frame] float32


def load_wav(path_or_bytes) -> WAVData:
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    if len(data) < 12 or data[0:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise WAVError("WAV file read error: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None or len(fmt) < 16:
        raise WAVError("WAV file read error: missing fmt or data chunk")
    tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE and len(fmt) >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID's first two bytes
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if channels <= 0:
        raise WAVError(f"Invalid channel count: {channels}. WAV file must have at least 1 channel.")  # :41-43
    frames = len(pcm) // block_align if block_align else 0
    if frames <= 0:
        raise WAVError("WAV file is empty (0 frames)")  # :45-47
    n = frames * channels
    if tag == 3 and bits == 32:
        x = np.frombuffer(pcm, "<f4", n).astype(np.float32)
    elif tag == 3 and bits == 64:
        x = np.frombuffer(pcm, "<f8", n).astype(np.float32)
    elif tag == 1 and bits == 16:
        x = (np.frombuffer(pcm, "<i2", n).astype(np.float32) / np.float32(32768.0))  # :78
    elif tag == 1 and bits == 32:
        x = (np.frombuffer(pcm, "<i4", n).astype(np.float32) / np.float32(2147483648.0))  # :86
    elif tag == 1 and bits == 24:
        raw = np.frombuffer(pcm, np.uint8, n * 3).reshape(-1, 3).astype(np.int32)
        v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = (v.astype(np.float32) / np.float32(8388608.0))
    else:
        raise WAVError("Unsupported WAV format")  # :89-91
    audio = np.ascontiguousarray(x.reshape(frames, channels).T)
    return WAVData(float(rate), channels, frames, audio)


# ------------------------------------------------------------------------------------------------
# InputLayout (VirtualSpeaker.swift:59-100) and HRIRChannelMap (:103-347)
# ------------------------------------------------------------------------------------------------
class InputLayout:
    def __init__(self, channels, name):
        self.channels = list(channels)
        self.name = name

    @staticmethod
    def detect(channelCount: int) -> "InputLayout":  # :89-99
        table = {2: InputLayout.stereo, 6: InputLayout.surround51, 8: InputLayout.surround71, 12: InputLayout.atmos714}
        if channelCount in table:
            return table[channelCount]
        return InputLayout([f"custom:Ch{i}" for i in range(channelCount)], f"{channelCount} Channel")


InputLayout.stereo = InputLayout(["FL", "FR"], "Stereo")  # :64-67
InputLayout.surround51 = InputLayout(["FL", "FR", "FC", "LFE", "BL", "BR"], "5.1 Surround")  # :70-73
InputLayout.surround71 = InputLayout(["FL", "FR", "FC", "LFE", "BL", "BR", "SL", "SR"], "7.1 Surround")  # :76-79
InputLayout.atmos714 = InputLayout(
    ["FL", "FR", "FC", "LFE", "BL", "BR", "SL", "SR", "TFL", "TFR", "TBL", "TBR"], "7.1.4 Atmos")  # :82-85

_LEFT_SIDE = {"FL", "BL", "SL", "TFL", "TBL", "FLC"}
_RIGHT_SIDE = {"FR", "BR", "SR", "TFR", "TBR", "FRC"}


class HRIRChannelMap:
    def __init__(self):
        self.mapping = {}

    def setMapping(self, speaker, leftEarIndex, rightEarIndex):  # :109-111
        self.mapping[speaker] = (leftEarIndex, rightEarIndex)

    def getIndices(self, speaker):  # :114-116
        return self.mapping.get(speaker)

    @staticmethod
    def interleavedPairs(speakers):  # :126-160
        m = HRIRChannelMap()
        for index, sp in enumerate(speakers):
            base = index * 2
            if sp in _RIGHT_SIDE:
                m.setMapping(sp, base + 1, base)
            else:
                m.setMapping(sp, base, base + 1)
        return m

    @staticmethod
    def splitBlocks(speakers):  # :201-210
        m = HRIRChannelMap()
        n = len(speakers)
        for index, sp in enumerate(speakers):
            m.setMapping(sp, index, index + n)
        return m

    @staticmethod
    def hesuvi7Channel(speakers):  # :224-250
        table = {"FL": (0, 1), "FR": (1, 0), "FC": (2, 2), "LFE": (2, 2), "BL": (3, 4), "BR": (4, 3),
                 "SL": (5, 6), "SR": (6, 5)}
        m = HRIRChannelMap()
        for sp in speakers:
            if sp in table:
                m.setMapping(sp, *table[sp])
        return m

    @staticmethod
    def hesuvi14Channel(speakers):  # :270-297
        table = {"FL": (0, 1), "FR": (8, 7), "FC": (6, 13), "LFE": (6, 13), "BL": (4, 5), "BR": (12, 11),
                 "SL": (2, 3), "SR": (10, 9)}
        m = HRIRChannelMap()
        for sp in speakers:
            if sp in table:
                m.setMapping(sp, *table[sp])
        return m

    @staticmethod
    def parseHeSuViFormat(text: str):  # :301-346
        names = {"FL": "FL", "L": "FL", "FR": "FR", "R": "FR", "FC": "FC", "C": "FC", "LFE": "LFE", "SUB": "LFE",
                 "BL": "BL", "RL": "BL", "BR": "BR", "RR": "BR", "SL": "SL", "SR": "SR", "TFL": "TFL",
                 "TFR": "TFR", "TBL": "TBL", "TBR": "TBR"}
        m = HRIRChannelMap()
        for line in re.split(r"\r\n|\n|\r", text):
            t = line.strip(" \t")
            if not t or t.startswith("#") or t.startswith(";"):
                continue
            parts = t.split("=")
            if len(parts) != 2:
                continue
            name = parts[0].strip(" \t")
            idx = []
            for tok in parts[1].strip(" \t").split(","):
                tok = tok.strip(" \t")
                if re.fullmatch(r"[+-]?\d+", tok):
                    idx.append(int(tok))
            if len(idx) != 2:
                continue
            m.setMapping(names.get(name.upper(), f"custom:{name}"), idx[0], idx[1])
        return m


# ------------------------------------------------------------------------------------------------
# HRIRManager.activatePreset build loop  (HRIRManager.swift:347-423)
# ------------------------------------------------------------------------------------------------
class HRIRError(Exception):
    pass


def _speaker_irs(wav: WAVData, targetSampleRate: float, inputLayout: InputLayout, hrirMap=None):
    if hrirMap is None:
        hrirMap = (HRIRChannelMap.hesuvi7Channel(inputLayout.channels) if wav.channelCount == 7
                   else HRIRChannelMap.hesuvi14Channel(inputLayout.channels))  # :355-360
    out = []
    for speaker in inputLayout.channels:
        idx = hrirMap.getIndices(speaker)
        if idx is None:
            continue  # :370-372
        l, r = idx
        if not (l < wav.channelCount and r < wav.channelCount):  # :375-379
            raise HRIRError(f"HRIR indices ({l}, {r}) out of range for {wav.channelCount} channels")
        left, right = wav.audioData[l], wav.audioData[r]
        if abs(wav.sampleRate - targetSampleRate) > 0.01:  # :389-403
            left = _b.resample_high_quality(left, wav.sampleRate, targetSampleRate)
            right = _b.resample_high_quality(right, wav.sampleRate, targetSampleRate)
        out.append((speaker, np.ascontiguousarray(left, np.float32), np.ascontiguousarray(right, np.float32)))
    if not out:
        raise HRIRError("No valid renderers created")  # :420-422
    return out


def activate_preset(wav: WAVData, targetSampleRate: float, inputLayout: InputLayout, blockSize: int = 512,
                    hrirMap=None):
    """Returns [VirtualSpeakerRenderer] exactly as the reference builds them (two engines per speaker)."""
    return [_b.VirtualSpeakerRenderer(sp, _b.ConvolutionEngine(l, blockSize), _b.ConvolutionEngine(r, blockSize))
            for sp, l, r in _speaker_irs(wav, targetSampleRate, inputLayout, hrirMap)]


def hrir_matrix(wav: WAVData, targetSampleRate: float, inputLayout: InputLayout, hrirMap=None) -> np.ndarray:
    """[S][2][taps] float32 impulse responses after mapping (+ resampling): input of direct_conv_f64."""
    irs = _speaker_irs(wav, targetSampleRate, inputLayout, hrirMap)
    taps = max(len(l) for _, l, _ in irs)
    h = np.zeros((len(irs), 2, taps), np.float32)
    for i, (_, l, r) in enumerate(irs):
        h[i, 0, : len(l)] = l
        h[i, 1, : len(r)] = r
    return h


# ------------------------------------------------------------------------------------------------
# EqualizerAPOParser.parse  (EqualizerAPOParser.swift:36-151)
# ------------------------------------------------------------------------------------------------
class EqualizerParseError(Exception):
    def __init__(self, filename, issues):
        self.filename = filename
        self.issues = issues  # [(lineNumber or None, reason)]
        details = "; ".join((f"line {ln}: {r}" if ln is not None else r) for ln, r in issues)
        super().__init__(f"Could not read {filename}: {details}")  # :12-20


_PREAMP_RE = re.compile(r"^Preamp\s*:\s*(\S+)\s+dB$", re.IGNORECASE)  # :27-30
_FILTER_RE = re.compile(
    r"^Filter(?:\s+([0-9]+))?\s*:\s+(ON|OFF)\s+(PK|LSC|HSC)\s+Fc\s+(\S+)\s+Hz\s+Gain\s+(\S+)\s+dB\s+Q\s+(\S+)$",
    re.IGNORECASE)  # :31-34
_SWIFT_DOUBLE_RE = re.compile(r"^[+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|0[xX][0-9a-fA-F.]+(?:[pP][+-]?\d+)?|inf|infinity|nan)$",
                              re.IGNORECASE)


def _finite_double(text):  # :153-156 (Swift Double(String): no surrounding whitespace, no underscores)
    if not _SWIFT_DOUBLE_RE.match(text):
        return None
    try:
        v = float.fromhex(text) if text.lower().lstrip("+-").startswith("0x") else float(text)
    except ValueError:
        return None
    return v if math.isfinite(v) else None


def parse_equalizer_apo(data: bytes, filename: str) -> dict:
    """Returns dict(preampDB, filters=[dict(sourceLine, sourceNumber, isEnabled, type, frequencyHz, gainDB, q)])."""
    if len(data) > 1_048_576:  # :37-42
        raise EqualizerParseError(filename, [(None, "file exceeds the 1 MiB limit")])
    try:
        source = data.decode("utf-8")
    except UnicodeDecodeError:
        raise EqualizerParseError(filename, [(None, "file is not valid UTF-8")])  # :43-48
    if source.startswith("\ufeff"):
        source = source[1:]  # :49-51
    preampDB, hasPreamp, declCount, filters, issues = 0.0, False, 0, [], []
    # components(separatedBy: .newlines): every newline character splits (CRLF gives an empty line between)
    for index, raw in enumerate(re.split("[\n\r\x0b\x0c\x85\u2028\u2029]", source)):
        lineNumber = index + 1
        line = raw.strip()
        if not line or line.startswith("#"):
            continue  # :62
        m = _PREAMP_RE.match(line)
        if m:  # :64-76
            if hasPreamp:
                issues.append((lineNumber, "duplicate Preamp directive"))
                continue
            v = _finite_double(m.group(1))
            if v is None:
                issues.append((lineNumber, "Preamp must be a finite number"))
                continue
            preampDB, hasPreamp = v, True
            continue
        if line.lower().startswith("filter"):  # :78-136
            declCount += 1
            if declCount > 64:
                issues.append((lineNumber, "more than 64 filter declarations are not allowed"))
                continue
            m = _FILTER_RE.match(line)
            if not m:
                issues.append((lineNumber, "malformed Filter directive"))
                continue
            num, onoff, ftype, fc, gain, q = m.groups()
            sourceNumber = int(num) if num else None
            isEnabled = onoff.upper() == "ON"
            type_ = {"PK": "peaking", "LSC": "lowShelf", "HSC": "highShelf"}[ftype.upper()]
            frequencyHz, gainDB, qv = _finite_double(fc), _finite_double(gain), _finite_double(q)
            numeric = []
            if frequencyHz is not None:
                if frequencyHz <= 0:
                    numeric.append("frequency must be positive")
            else:
                numeric.append("frequency must be a finite number")
            if gainDB is None:
                numeric.append("gain must be a finite number")
            if qv is not None:
                if qv <= 0:
                    numeric.append("Q must be positive")
            else:
                numeric.append("Q must be a finite number")
            if numeric:
                issues.extend((lineNumber, r) for r in numeric)
                continue
            filters.append(dict(sourceLine=lineNumber, sourceNumber=sourceNumber, isEnabled=isEnabled, type=type_,
                                frequencyHz=frequencyHz, gainDB=gainDB, q=qv))
            continue
        if line.lower().startswith("preamp"):  # :138-142
            issues.append((lineNumber, "malformed Preamp directive"))
        else:
            issues.append((lineNumber, "unsupported directive"))
    if not issues and preampDB == 0 and not any(f["isEnabled"] for f in filters):  # :145-147
        issues.append((None, "effective configuration must contain a non-zero preamp or an enabled supported filter"))
    if issues:
        raise EqualizerParseError(filename, issues)
    return dict(preampDB=preampDB, filters=filters)


# ------------------------------------------------------------------------------------------------
# AudioEffectGraph.process routing  (AudioEffectGraph.swift:179-246)
# ------------------------------------------------------------------------------------------------
class AudioEffectGraphModel:
    """spatial / equalizer: objects with process(left, right_or_None) -> (L, R); spatial has isReady."""

    def __init__(self, spatial, equalizer, maxFramesPerCallback: int = 4096):
        assert 0 < maxFramesPerCallback <= 4096  # :81
        self.spatial, self.equalizer, self.maxFramesPerCallback = spatial, equalizer, maxFramesPerCallback
        self.equalizerActive = False

    def process(self, inputLeft, inputRight=None):
        left = np.asarray(inputLeft, np.float32)
        right = None if inputRight is None else np.asarray(inputRight, np.float32)
        if len(left) <= 0:
            return left, left
        assert len(left) <= self.maxFramesPerCallback  # :187
        if self.spatial is not None and self.spatial.isReady:  # :195-222
            l, r = self.spatial.process(left, right)
            if self.equalizerActive:
                l, r = self.equalizer.process(l, r)
            return l, r
        l, r = left.copy(), (left.copy() if right is None else right.copy())  # :225-229, 241-245
        if self.equalizerActive:
            l, r = self.equalizer.process(l, r)  # :230-236
        return l, r
