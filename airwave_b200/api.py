"""Host-side mirror of the reference's interface for the binaural path, over the C ABI.

Names, argument meaning and error behaviour follow the reference's Swift types so parity tests
read like the reference's XCTest files:

    WAVLoader / WAVData            Airwave/WAVLoader.swift
    InputLayout / HRIRChannelMap   Airwave/VirtualSpeaker.swift
    ConvolutionEngine              Airwave/ConvolutionEngine.swift
    VirtualSpeakerRenderer         Airwave/HRIRManager.swift:84-88
    RealtimeAudioProcessor         Airwave/RealtimeAudioProcessor.swift
    ParametricEqualizerState/Processor, BiquadCoefficientBuilder, EqualizerAPOParser
    AudioEffectGraph               Airwave/AudioEffectGraph.swift

plus the batched objects that exist only here: HRIRBank (one frequency-domain filter bank per
preset x rate x block, resident in HBM) and BinauralEngine (n streams in lock-step).

Every arithmetic step runs in libairwave_cuda.so on the GPU; nothing here computes audio.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import AirwaveError, EqFilter, EngineConfig

__all__ = [
    "AirwaveError", "WAVData", "WAVLoader", "InputLayout", "HRIRChannelMap", "HRIRBank", "BinauralEngine",
    "ConvolutionEngine", "VirtualSpeakerRenderer", "RealtimeAudioProcessor", "BiquadCoefficientBuilder",
    "BiquadCoefficientError", "EqualizerAPOParser", "EqualizerParseError", "ParametricEqualizerState",
    "ParametricEqualizerProcessor", "ParametricEqualizerPreparationError", "AudioEffectGraph",
    "AudioEffectPreparationResult", "Resampler", "FFTSetupManager", "device_count", "PinnedBuffer",
]

SPEAKERS = ["FL", "FR", "FC", "LFE", "BL", "BR", "SL", "SR", "TFL", "TFR", "TBL", "TBR", "FLC", "FRC", "BC"]


def device_count() -> int:
    return L.lib().aw_device_count()


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _iarr(values: Sequence[int]):
    return (C.c_int * max(len(values), 1))(*values)


# ------------------------------------------------------------------------------------------------
class FFTSetupManager:
    """FFTSetupManager.shared (FFTSetupManager.swift:13-69): the per-device twiddle/plan cache."""

    @staticmethod
    def getSetup(log2n: int, device: int = 0) -> bool:
        L.check(L.lib().aw_plan_prepare(device, log2n))
        return True

    @staticmethod
    def getCacheStats(device: int = 0):
        count = C.c_int()
        sizes = (C.c_int * 32)()
        L.check(L.lib().aw_plan_cache_stats(device, C.byref(count), sizes, 32))
        return count.value, sorted(sizes[i] for i in range(min(count.value, 32)))


class WAVData:
    """WAVLoader.swift:12-17."""

    def __init__(self, handle):
        self._h = handle
        sr, ch, fr = C.c_double(), C.c_int(), C.c_int()
        L.check(L.lib().aw_wav_info(handle, C.byref(sr), C.byref(ch), C.byref(fr)))
        self.sampleRate, self.channelCount, self.frameCount = sr.value, ch.value, fr.value

    @property
    def audioData(self) -> np.ndarray:
        out = np.empty((self.channelCount, self.frameCount), np.float32)
        for c in range(self.channelCount):
            out[c] = np.ctypeslib.as_array(L.lib().aw_wav_channel(self._h, c), (self.frameCount,))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib().aw_wav_destroy(self._h)
            self._h = None


class WAVLoader:
    """WAVLoader.load (WAVLoader.swift:26-99)."""

    @staticmethod
    def load(source) -> WAVData:
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray)):
            L.check(L.lib().aw_wav_load_memory(bytes(source), len(source), C.byref(h)))
        else:
            L.check(L.lib().aw_wav_load(str(source).encode(), C.byref(h)))
        return WAVData(h)


class InputLayout:
    """VirtualSpeaker.swift:59-100."""

    def __init__(self, channels, name, code=0):
        self.channels, self.name, self.code = list(channels), name, code

    @staticmethod
    def _from_code(code: int, name: str) -> "InputLayout":
        buf = (C.c_int * 16)()
        n = L.lib().aw_layout_speakers(code, buf, 16)
        return InputLayout([SPEAKERS[buf[i]] for i in range(n)], name, code)

    @staticmethod
    def detect(channelCount: int) -> "InputLayout":
        table = {2: "stereo", 6: "surround51", 8: "surround71", 12: "atmos714"}
        if channelCount in table:
            return getattr(InputLayout, table[channelCount])()
        return InputLayout([f"custom:Ch{i}" for i in range(channelCount)], f"{channelCount} Channel")

    @staticmethod
    def stereo(): return InputLayout._from_code(L.LAYOUT_STEREO, "Stereo")

    @staticmethod
    def surround51(): return InputLayout._from_code(L.LAYOUT_SURROUND51, "5.1 Surround")

    @staticmethod
    def surround71(): return InputLayout._from_code(L.LAYOUT_SURROUND71, "7.1 Surround")

    @staticmethod
    def atmos714(): return InputLayout._from_code(L.LAYOUT_ATMOS714, "7.1.4 Atmos")


class HRIRChannelMap:
    """VirtualSpeaker.swift:103-347 (HeSuVi maps and the text parser run in the C library)."""

    def __init__(self, mapping=None):
        self.mapping = dict(mapping or {})

    def getIndices(self, speaker):
        return self.mapping.get(speaker)

    @staticmethod
    def _hesuvi(wav_channels: int, speakers) -> "HRIRChannelMap":
        known = [s for s in speakers if s in SPEAKERS]
        codes = _iarr([SPEAKERS.index(s) for s in known])
        l, r = (C.c_int * max(len(known), 1))(), (C.c_int * max(len(known), 1))()
        L.check(L.lib().aw_hesuvi_map(wav_channels, codes, len(known), l, r))
        return HRIRChannelMap({s: (l[i], r[i]) for i, s in enumerate(known) if l[i] >= 0})

    @staticmethod
    def hesuvi14Channel(speakers): return HRIRChannelMap._hesuvi(14, speakers)

    @staticmethod
    def hesuvi7Channel(speakers): return HRIRChannelMap._hesuvi(7, speakers)

    @staticmethod
    def parseHeSuViFormat(text: str) -> "HRIRChannelMap":
        l, r = (C.c_int * len(SPEAKERS))(), (C.c_int * len(SPEAKERS))()
        L.check(L.lib().aw_hesuvi_parse(text.encode(), l, r))
        return HRIRChannelMap({SPEAKERS[i]: (l[i], r[i]) for i in range(len(SPEAKERS)) if l[i] >= 0})


class Resampler:
    """Resampler.resampleHighQuality (Resampler.swift:31-68), computed on the device.

    ``correct=False`` (default) reproduces the reference literally, including its time compression of up-sampled responses
     and the refusal of down-sampling; ``correct=True`` is the flagged alternative: time-correct linear
    interpolation in float64, down-sampling allowed."""

    @staticmethod
    def resampleHighQuality(input, fromRate: float, toRate: float, device: int = 0, correct: bool = False) -> np.ndarray:
        x = _f32(input)
        n = L.lib().aw_resample_output_count(len(x), fromRate, toRate)
        out = np.zeros(max(n, len(x), 1), np.float32)
        written = C.c_int()
        L.check(L.lib().aw_resample_ex(device, x.ctypes.data_as(C.POINTER(C.c_float)), len(x), fromRate, toRate,
                                       L.RESAMPLE_CORRECT if correct else L.RESAMPLE_REFERENCE,
                                       out.ctypes.data_as(C.POINTER(C.c_float)), len(out), C.byref(written)))
        return out[: written.value]


# ------------------------------------------------------------------------------------------------
class HRIRBank:
    """Frequency-domain HRIR filter bank resident in HBM (aw_bank): what HRIRManager.activatePreset's
    build loop + ConvolutionEngine.init produce, once per (preset, rate, block) instead of per engine."""

    def __init__(self, pcm, src_rate: float, dst_rate: float, left_idx, right_idx, block: int, device: int = 0,
                 correct_resampling: bool = False):
        pcm = _f32(pcm)
        assert pcm.ndim == 2
        h = C.c_void_p()
        L.check(L.lib().aw_bank_create_ex(device, pcm.ctypes.data_as(C.POINTER(C.c_float)), pcm.shape[0], pcm.shape[1],
                                          src_rate, dst_rate, _iarr(list(left_idx)), _iarr(list(right_idx)), len(left_idx),
                                          block, L.RESAMPLE_CORRECT if correct_resampling else L.RESAMPLE_REFERENCE, C.byref(h)))
        self._h, self.device = h, device
        self._read_info()

    @classmethod
    def from_wav(cls, wav: WAVData, dst_rate: float, layout: InputLayout, block: int, device: int = 0) -> "HRIRBank":
        self = cls.__new__(cls)
        h = C.c_void_p()
        L.check(L.lib().aw_bank_create_from_wav(device, wav._h, dst_rate, layout.code, block, C.byref(h)))
        self._h, self.device = h, device
        self._read_info()
        return self

    def _read_info(self):
        s, b, p, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib().aw_bank_info(self._h, C.byref(s), C.byref(b), C.byref(p), C.byref(t)))
        self.n_speakers, self.block, self.partitions, self.taps = s.value, b.value, p.value, t.value
        self.rows = L.lib().aw_bank_rows(self._h)   # distinct filter pairs (FC and LFE share one): FDL rows per stream

    def read(self):
        spec = np.zeros((self.n_speakers, self.partitions, self.block, 4), np.float32)
        ny = np.zeros((self.n_speakers, self.partitions, 2), np.float32)
        L.check(L.lib().aw_bank_read(self._h, spec.ctypes.data_as(C.POINTER(C.c_float)), ny.ctypes.data_as(C.POINTER(C.c_float))))
        return spec, ny

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib().aw_bank_destroy(self._h)
            self._h = None


def _pack_definition(definition):
    """definition: None | dict(preampDB=, filters=[dict(type, frequencyHz, gainDB, q, isEnabled, sourceLine)])."""
    if definition is None:
        return 0.0, None, -1
    filters = definition.get("filters", [])
    arr = (EqFilter * max(len(filters), 1))()
    for i, f in enumerate(filters):
        arr[i] = EqFilter(L.FILTER_TYPES[f["type"]], 1 if f.get("isEnabled", True) else 0, f["frequencyHz"], f["gainDB"],
                          f["q"], int(f.get("sourceLine", 0) or 0), -1 if f.get("sourceNumber") is None else int(f["sourceNumber"]))
    return float(definition.get("preampDB", 0.0)), arr, len(filters)


class PinnedBuffer:
    """Page-locked host array (cudaHostAlloc) for the pipelined host path."""

    def __init__(self, shape, dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = L.lib().aw_host_alloc(max(self.nbytes, 1))
        if not self._p:
            raise AirwaveError(L.ERR_OUT_OF_MEMORY, "cudaHostAlloc failed")
        buf = (C.c_char * self.nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        if getattr(self, "_p", None):
            self.array = None
            L.lib().aw_host_free(self._p)
            self._p = None


class BinauralEngine:
    """aw_engine: n streams x (RealtimeAudioProcessor + per-speaker convolvers + EQ) rendered in lock-step."""

    def __init__(self, n_streams: int, n_speakers: int, block: int = 512, sample_rate: float = 48000.0,
                 max_frames_per_call: int = 4096, max_partitions: int = 0, device: int = 0, literal_stereo: bool = False,
                 pipelined: bool = False, overlap_eq: bool = False):
        cfg = EngineConfig(device, n_streams, n_speakers, block, sample_rate, max_frames_per_call, max_partitions,
                           (L.ENGINE_LITERAL_STEREO if literal_stereo else 0) | (L.ENGINE_PIPELINED if pipelined else 0)
                           | (L.ENGINE_OVERLAP_EQ if overlap_eq else 0))
        h = C.c_void_p()
        L.check(L.lib().aw_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.n_streams, self.n_speakers, self.block, self.sample_rate = n_streams, n_speakers, block, sample_rate
        self.max_frames_per_call, self.device = max_frames_per_call, device
        self._banks = {}

    # -- control -------------------------------------------------------------------------------
    def set_bank(self, bank: Optional[HRIRBank], first: int = 0, count: Optional[int] = None) -> None:
        count = self.n_streams - first if count is None else count
        L.check(L.lib().aw_engine_set_bank(self._h, first, count, bank._h if bank is not None else None))
        self._banks[(first, count)] = bank   # keep the bank alive while the engine references it

    def _eq(self, fn, definition, first, count, *extra):
        count = self.n_streams - first if count is None else count
        preamp, arr, n = _pack_definition(definition)
        bi, br = C.c_int(-1), C.c_int(0)
        status = fn(self._h, first, count, preamp, arr, n, *extra, C.byref(bi), C.byref(br))
        L.check(status, bi.value, br.value)

    def eq_prepare(self, definition, first=0, count=None): self._eq(L.lib().aw_engine_eq_prepare, definition, first, count)
    def eq_update(self, definition, first=0, count=None): self._eq(L.lib().aw_engine_eq_update, definition, first, count)
    def eq_install_state(self, definition, first=0, count=None): self._eq(L.lib().aw_engine_eq_install_state, definition, first, count)

    def eq_set_target(self, definition, first=0, count=None, drain_retired=False):
        self._eq(L.lib().aw_engine_eq_set_target, definition, first, count, 1 if drain_retired else 0)

    def eq_drain_retired(self, first=0, count=None):
        L.check(L.lib().aw_engine_eq_drain_retired(self._h, first, self.n_streams - first if count is None else count))

    def eq_active(self, active: bool, first=0, count=None):
        L.check(L.lib().aw_engine_eq_active(self._h, first, self.n_streams - first if count is None else count, 1 if active else 0))

    def eq_hold_publication(self, held: bool, first=0, count=None):
        L.check(L.lib().aw_engine_eq_hold_publication(self._h, first, self.n_streams - first if count is None else count, 1 if held else 0))

    def reset(self, first=0, count=None, spatial=True, eq=False):
        what = (L.RESET_SPATIAL if spatial else 0) | (L.RESET_EQ if eq else 0)
        L.check(L.lib().aw_engine_reset(self._h, first, self.n_streams - first if count is None else count, what))

    # -- render --------------------------------------------------------------------------------
    def process(self, x) -> np.ndarray:
        """x: [n_streams][n_speakers][frames] float32 (host) -> [n_streams][2][frames]."""
        x = _f32(x)
        assert x.shape[:2] == (self.n_streams, self.n_speakers), x.shape
        frames = x.shape[2]
        out = np.full((self.n_streams, 2, frames), np.nan, np.float32)
        L.check(L.lib().aw_engine_process(self._h, _ptr(x), _ptr(out), frames))
        return out

    def process_stereo(self, left, right=None, alias_outputs: bool = False):
        l = _f32(left)
        r = None if right is None else _f32(right)
        n = len(l)
        outL = np.full(n, np.nan, np.float32)
        outR = outL if alias_outputs else np.full(n, np.nan, np.float32)
        L.check(L.lib().aw_engine_process_stereo(self._h, _ptr(l), None if r is None else _ptr(r), _ptr(outL), _ptr(outR), n))
        return outL, outR

    def process_device(self, in_ptr: int, in_ss: int, in_cs: int, out_ptr: int, out_ss: int, out_cs: int, frames: int) -> None:
        L.check(L.lib().aw_engine_process_device(self._h, in_ptr, in_ss, in_cs, out_ptr, out_ss, out_cs, frames))

    def submit(self, in_ptr: int, out_ptr: int, frames: int) -> None:
        L.check(L.lib().aw_engine_submit(self._h, in_ptr, out_ptr, frames))

    def submit_device(self, in_ptr: int, in_ss: int, in_cs: int, out_ptr: int, frames: int) -> None:
        """Device-resident input, host output copied asynchronously (valid after wait())."""
        L.check(L.lib().aw_engine_submit_device(self._h, in_ptr, in_ss, in_cs, out_ptr, frames))

    def wait(self) -> None:
        L.check(L.lib().aw_engine_wait(self._h))

    def flush(self) -> None:
        """Orders the engine's stream behind an overlapped equalizer still in flight (AW_ENGINE_OVERLAP_EQ)."""
        L.check(L.lib().aw_engine_flush(self._h))

    def counters(self) -> dict:
        v = [C.c_ulonglong() for _ in range(4)]
        L.check(L.lib().aw_engine_counters(self._h, *[C.byref(x) for x in v]))
        return dict(kernel_launches=v[0].value, blocks=v[1].value, h2d_bytes=v[2].value, d2h_bytes=v[3].value)

    def plan(self) -> dict:
        f, m, p = C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib().aw_engine_plan(self._h, C.byref(f), C.byref(m), C.byref(p)))
        return dict(fused_tile=f.value, mac_tile=m.value, partitions_cap=p.value, kernels=self.kernels(),
                    tensor_map_tma=bool(L.lib().aw_engine_uses_tensor_maps(self._h)))

    def kernels(self) -> list:
        """Names of the kernels launched per block, in launch order (e.g. ['k_persistent<8,4>'])."""
        buf = C.create_string_buffer(256)
        L.check(L.lib().aw_engine_kernels(self._h, buf, 256))
        return buf.value.decode().split(";")

    def profile_begin(self, max_blocks: int) -> None:
        L.check(L.lib().aw_engine_profile_begin(self._h, max_blocks))

    def profile_end(self) -> dict:
        """{kernel name: {ms, launches}} summed over the profiled blocks; 'k_eq' = the equalizer launches of a call."""
        ms = (C.c_double * 4)()
        cnt = (C.c_ulonglong * 4)()
        L.check(L.lib().aw_engine_profile_end(self._h, ms, cnt))
        out = {n: dict(ms=ms[i], launches=cnt[i]) for i, n in enumerate(self.kernels())}
        if cnt[3] and ms[3] > 0:
            out["k_eq"] = dict(ms=ms[3], launches=cnt[3])
        return out

    @property
    def cuda_stream(self) -> int:
        return L.lib().aw_engine_stream(self._h) or 0

    def close(self):
        if getattr(self, "_h", None):
            L.lib().aw_engine_destroy(self._h)
            self._h = None
            self._banks = {}

    def __del__(self):
        self.close()


# ------------------------------------------------------------------------------------------------
# Single-stream mirrors of the reference classes
# ------------------------------------------------------------------------------------------------
class ConvolutionEngine:
    """ConvolutionEngine (ConvolutionEngine.swift:14-408): mono in, mono out, one impulse response."""

    def __init__(self, hrirSamples, blockSize: int = 512, device: int = 0):
        self.hrirSamples = _f32(hrirSamples)
        self.blockSize = blockSize
        try:
            self._bank = HRIRBank(self.hrirSamples.reshape(1, -1) if len(self.hrirSamples) else np.zeros((1, 1), np.float32),
                                  48000.0, 48000.0, [0], [0], blockSize, device)
        except AirwaveError as e:
            if e.status == L.ERR_INVALID_BLOCK_SIZE:
                raise ValueError("ConvolutionEngine init failed (init? returned nil)") from e
            raise
        self._engine = BinauralEngine(1, 1, blockSize, 48000.0, blockSize, self._bank.partitions, device)
        self._engine.set_bank(self._bank)

    @property
    def partitionCount(self) -> int:
        return self._bank.partitions

    def process(self, input, frameCount: Optional[int] = None):
        count = self.blockSize if frameCount is None else frameCount
        if count != self.blockSize:
            return None   # ConvolutionEngine.swift:370-372: silently returns
        x = _f32(input)[: self.blockSize].reshape(1, 1, -1)
        return self._engine.process(x)[0, 0].copy()

    def processAndAccumulate(self, input, outputAccumulator: np.ndarray) -> None:
        outputAccumulator += self.process(input)   # ConvolutionEngine.swift:388-394

    def reset(self) -> None:
        self._engine.reset()


class VirtualSpeakerRenderer:
    def __init__(self, speaker, convolverLeftEar: ConvolutionEngine, convolverRightEar: ConvolutionEngine):
        self.speaker, self.convolverLeftEar, self.convolverRightEar = speaker, convolverLeftEar, convolverRightEar


class RealtimeAudioProcessor:
    """RealtimeAudioProcessor (RealtimeAudioProcessor.swift:11-191) for one stream.

    The renderers' impulse responses are gathered into one HRIRBank.  ``literalStereo`` keeps the
    reference rule (at most two renderers, fed by left/right, :145-147); ``False`` feeds renderer i
    from input channel i."""

    def __init__(self, renderers: Iterable[VirtualSpeakerRenderer], blockSize: int = 512, maxFramesPerCallback: int = 4096,
                 literalStereo: bool = True, device: int = 0):
        if blockSize <= 0 or maxFramesPerCallback <= 0:
            raise AssertionError("precondition failed")
        self.renderers = list(renderers)
        self.blockSize, self.maxFramesPerCallback, self.literalStereo = blockSize, maxFramesPerCallback, literalStereo
        n = len(self.renderers)
        self.inputCount = 2 if literalStereo else max(n, 1)
        self._engine = BinauralEngine(1, self.inputCount, blockSize, 48000.0, maxFramesPerCallback, 0, device,
                                      literal_stereo=literalStereo)
        self._bank = None
        if n:
            taps = max(max(len(r.convolverLeftEar.hrirSamples), len(r.convolverRightEar.hrirSamples)) for r in self.renderers)
            pcm = np.zeros((2 * n, max(taps, 1)), np.float32)
            for i, r in enumerate(self.renderers):
                pcm[2 * i, : len(r.convolverLeftEar.hrirSamples)] = r.convolverLeftEar.hrirSamples
                pcm[2 * i + 1, : len(r.convolverRightEar.hrirSamples)] = r.convolverRightEar.hrirSamples
            use = min(n, 2) if literalStereo else n
            self._bank = HRIRBank(pcm, 48000.0, 48000.0, [2 * i for i in range(use)], [2 * i + 1 for i in range(use)],
                                  blockSize, device)
            self._engine.set_bank(self._bank)

    def process(self, inputLeft, inputRight=None, frameCount: Optional[int] = None, aliasOutputs: bool = False):
        if self.inputCount != 2:
            raise ValueError("use process_channels for non-stereo input")
        l = _f32(inputLeft)
        n = len(l) if frameCount is None else frameCount
        if n > self.maxFramesPerCallback:
            raise AssertionError("precondition(frameCount <= maxFramesPerCallback)")
        try:
            return self._engine.process_stereo(l[:n], None if inputRight is None else _f32(inputRight)[:n], aliasOutputs)
        except AirwaveError as e:
            if e.status == L.ERR_FRAME_COUNT:
                raise AssertionError("precondition(frameCount <= maxFramesPerCallback)") from e
            raise

    def process_channels(self, inputs):
        x = np.stack([_f32(a) for a in inputs])[None]
        out = self._engine.process(x)
        return out[0, 0].copy(), out[0, 1].copy()

    def reset(self) -> None:
        self._engine.reset()


class BiquadCoefficientError(Exception):
    NAMES = {1: "invalidSampleRate", 2: "invalidFrequency", 3: "invalidQ", 4: "nonFiniteInput", 5: "nonFiniteCoefficients"}

    def __init__(self, code: int):
        super().__init__(self.NAMES.get(code, str(code)))
        self.code, self.name = code, self.NAMES.get(code, str(code))


class BiquadCoefficientBuilder:
    @staticmethod
    def make(type, gainDB: float, frequencyHz: float, q: float, sampleRate: float) -> np.ndarray:
        out = np.zeros(5, np.float64)
        rc = L.lib().aw_biquad_make(L.FILTER_TYPES[type], gainDB, frequencyHz, q, sampleRate, out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise BiquadCoefficientError(rc)
        return out


class EqualizerParseError(Exception):
    def __init__(self, filename: str, issues: str):
        super().__init__(f"Could not read {filename}: {issues}")
        self.filename, self.issues_text = filename, issues
        self.issues = []
        for part in issues.split("; "):
            if part.startswith("line "):
                head, _, reason = part.partition(": ")
                self.issues.append((int(head[5:]), reason))
            else:
                self.issues.append((None, part))


class EqualizerAPOParser:
    maximumDataSize = 1_048_576
    maximumFilterCount = 64

    @staticmethod
    def parse(data: bytes, filename: str) -> dict:
        filters = (EqFilter * 64)()
        n, preamp = C.c_int(), C.c_double()
        issues = C.create_string_buffer(16384)
        rc = L.lib().aw_eq_parse(bytes(data), len(data), C.byref(preamp), filters, 64, C.byref(n), issues, len(issues))
        if rc == L.ERR_EQ_PARSE:
            raise EqualizerParseError(filename, issues.value.decode("utf-8", "replace"))
        L.check(rc)
        return dict(preampDB=preamp.value, filters=[
            dict(sourceLine=f.source_line, sourceNumber=None if f.source_number < 0 else f.source_number,
                 isEnabled=bool(f.enabled), type=L.FILTER_NAMES[f.type], frequencyHz=f.frequency_hz, gainDB=f.gain_db, q=f.q)
            for f in filters[: n.value]])


class ParametricEqualizerPreparationError(Exception):
    NAMES = {L.ERR_EQ_INVALID_SAMPLE_RATE: "invalidSampleRate", L.ERR_EQ_NON_FINITE_PREAMP: "nonFinitePreamp",
             L.ERR_EQ_TOO_MANY_FILTERS: "tooManyFilters", L.ERR_EQ_INVALID_FILTER: "invalidFilter"}

    def __init__(self, err: AirwaveError):
        super().__init__(err.message)
        self.status, self.index, self.filter_error = err.status, err.bad_index, err.bad_reason
        self.name = self.NAMES.get(err.status, str(err.status))


def _eq_guard(fn, *args, **kw):
    try:
        return fn(*args, **kw)
    except AirwaveError as e:
        if e.status in ParametricEqualizerPreparationError.NAMES:
            raise ParametricEqualizerPreparationError(e) from e
        raise


class ParametricEqualizerState:
    """ParametricEqualizerState (ParametricEqualizerProcessor.swift:16-98) used directly: no crossfade."""

    def __init__(self, definition, sampleRate: float, maxFrames: int = 4096, device: int = 0):
        if not (np.isfinite(sampleRate) and sampleRate > 0):
            raise ParametricEqualizerPreparationError(AirwaveError(L.ERR_EQ_INVALID_SAMPLE_RATE, "Sample rate must be finite and positive."))
        self.sampleRate = sampleRate
        self._engine = BinauralEngine(1, 2, 512, sampleRate, maxFrames, 0, device)
        _eq_guard(self._engine.eq_install_state, definition)
        self._maxFrames = maxFrames

    def process(self, left, right=None):
        l = _f32(left)
        r = None if right is None else _f32(right)
        outL, outR = np.empty(len(l), np.float32), np.empty(len(l), np.float32)
        for a in range(0, len(l), self._maxFrames):
            b = min(len(l), a + self._maxFrames)
            outL[a:b], outR[a:b] = self._engine.process_stereo(l[a:b], None if r is None else r[a:b])
        return outL, outR


class ParametricEqualizerProcessor:
    """ParametricEqualizerProcessor (ParametricEqualizerProcessor.swift:121-408) for one stereo stream."""

    @staticmethod
    def prepare(definition, sampleRate: float) -> ParametricEqualizerState:
        return ParametricEqualizerState(definition, sampleRate)

    def __init__(self, sampleRate: float, maxFramesPerCallback: int = 4096, device: int = 0):
        if not (np.isfinite(sampleRate) and sampleRate > 0):
            raise ParametricEqualizerPreparationError(AirwaveError(L.ERR_EQ_INVALID_SAMPLE_RATE, "Sample rate must be finite and positive."))
        if not (0 < maxFramesPerCallback <= 4096):
            raise ParametricEqualizerPreparationError(AirwaveError(L.ERR_EQ_TOO_MANY_FILTERS, "maxFramesPerCallback out of range"))
        self.sampleRate, self.maxFramesPerCallback = sampleRate, maxFramesPerCallback
        self._engine = BinauralEngine(1, 2, 512, sampleRate, maxFramesPerCallback, 0, device)
        # a processor exists and sits in the callback path; its initial state is unity (:158-159)
        self._engine.eq_install_state(None)

    def setTarget(self, definition) -> None:
        _eq_guard(self._engine.eq_set_target, definition, drain_retired=False)

    def reset(self) -> None:
        self._engine.reset(spatial=False, eq=True)

    def drainRetiredStates(self) -> None:
        self._engine.eq_drain_retired()

    def holdPublicationLock(self, held: bool) -> None:
        self._engine.eq_hold_publication(held)

    def process(self, left, right=None):
        l = _f32(left)
        if len(l) > self.maxFramesPerCallback:
            raise AssertionError("precondition(frameCount <= maxFramesPerCallback)")
        return self._engine.process_stereo(l, None if right is None else _f32(right))


class AudioEffectPreparationResult:
    def __init__(self, runnableEffects, equalizerWarning=None):
        self.runnableEffects, self.equalizerWarning = set(runnableEffects), equalizerWarning

    @property
    def noEffectCanRun(self) -> bool:
        return not self.runnableEffects


class AudioEffectGraph:
    """AudioEffectGraph (AudioEffectGraph.swift:65-248) for one stereo stream, with the production effects
    (HRIRManager as spatial, EqualizerRuntimeEffect as equalizer) living inside one aw_engine."""

    maximumCallbackFrames = 4096

    def __init__(self, maxFramesPerCallback: int = 4096, blockSize: int = 512, device: int = 0):
        assert 0 < maxFramesPerCallback <= self.maximumCallbackFrames   # AudioEffectGraph.swift:81
        self.maxFramesPerCallback, self.blockSize, self.device = maxFramesPerCallback, blockSize, device
        self._engine = None
        self._sampleRate = None
        self._bank = None
        self._wav = None

    def _ensure_engine(self, sampleRate: float):
        if self._engine is None or self._sampleRate != sampleRate:
            if not (np.isfinite(sampleRate) and sampleRate > 0):
                raise AirwaveError(L.ERR_EQ_INVALID_SAMPLE_RATE, "Output sample rate is invalid.")
            self._engine = BinauralEngine(1, 2, self.blockSize, sampleRate, self.maxFramesPerCallback, 0, self.device,
                                          literal_stereo=True)
            self._sampleRate = sampleRate
            if self._wav is not None:
                self._publish_bank()

    def _publish_bank(self):
        self._bank = HRIRBank.from_wav(self._wav, self._sampleRate, InputLayout.stereo(), self.blockSize, self.device)
        self._engine.set_bank(self._bank)

    # HRIRManager.activatePreset / deactivatePreset (HRIRManager.swift:316-475), inputLayout: .stereo as in production
    def activatePreset(self, wav: WAVData, targetSampleRate: float) -> None:
        self._wav = wav
        self._ensure_engine(targetSampleRate)
        self._publish_bank()

    def deactivatePreset(self) -> None:
        self._wav, self._bank = None, None
        if self._engine is not None:
            self._engine.set_bank(None)

    @property
    def spatialIsReady(self) -> bool:
        return self._bank is not None

    def _result(self, definition, call):
        runnable = {"spatial"} if self.spatialIsReady else set()
        try:
            call(definition)
            if definition is not None:
                runnable.add("equalizer")
            return AudioEffectPreparationResult(runnable)
        except AirwaveError as e:
            line = None
            if e.status == L.ERR_EQ_INVALID_FILTER and definition is not None:
                enabled = [f for f in definition.get("filters", []) if f.get("isEnabled", True)]
                if 0 <= e.bad_index < len(enabled):
                    line = enabled[e.bad_index].get("sourceLine")   # EqualizerRuntimeEffect.swift:85-89
            reason = e.message.split("is invalid: ", 1)[-1]
            return AudioEffectPreparationResult(runnable, dict(filterLine=line, reason=reason))

    def prepare(self, sampleRate: float, equalizerDefinition) -> AudioEffectPreparationResult:
        try:
            self._ensure_engine(sampleRate)
        except AirwaveError as e:
            return AudioEffectPreparationResult(set(), dict(filterLine=None, reason=e.message))
        return self._result(equalizerDefinition, self._engine.eq_prepare)

    def updateEqualizer(self, definition) -> AudioEffectPreparationResult:
        if self._engine is None:
            return AudioEffectPreparationResult(set(), dict(filterLine=None, reason="Equalizer has not been prepared for an output."))
        return self._result(definition, self._engine.eq_update)

    def process(self, inputLeft, inputRight=None):
        l = _f32(inputLeft)
        if len(l) <= 0:
            return l, l
        assert len(l) <= self.maxFramesPerCallback   # AudioEffectGraph.swift:187
        if self._engine is None:   # never prepared: neither effect can run -> passthrough
            self._ensure_engine(48000.0)
        return self._engine.process_stereo(l, None if inputRight is None else _f32(inputRight))
