"""ctypes binding of oracle/airwave_oracle.c (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Class and method names mirror the reference's Swift types so the parity tests read like the
reference's own XCTest files.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libairwave_oracle.so")

__all__ = [
    "build_oracle", "lib", "FFTSetup", "ConvolutionEngine", "VirtualSpeakerRenderer",
    "RealtimeAudioProcessor", "resample_high_quality", "resample_linear_f64", "resample_output_count",
    "biquad_make", "BiquadCoefficientError", "ParametricEqualizerState",
    "ParametricEqualizerProcessor", "ParametricEqualizerPreparationError",
    "direct_conv_f64", "synth_fill", "synth_block", "bench_render", "max_threads", "CpuBatch",
]


def build_oracle(force: bool = False) -> str:
    """Compile the C restatement (gcc) if it is missing or stale."""
    src = os.path.join(_HERE, "airwave_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_LIB_PATH)
        fp, dp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.or_fft_create.restype = vp; L.or_fft_create.argtypes = [C.c_int]
        L.or_fft_destroy.argtypes = [vp]
        L.or_fft_zrip.argtypes = [vp, fp, fp, C.c_int]
        L.or_conv_create.restype = vp; L.or_conv_create.argtypes = [fp, C.c_int, C.c_int, vp]
        L.or_conv_destroy.argtypes = [vp]
        L.or_conv_partition_count.argtypes = [vp]
        L.or_conv_process.argtypes = [vp, fp, fp]
        L.or_conv_process_array.argtypes = [vp, fp, fp, C.c_int]
        L.or_conv_process_and_accumulate.argtypes = [vp, fp, fp]
        L.or_conv_reset.argtypes = [vp]
        L.or_rap_create.restype = vp
        L.or_rap_create.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]
        L.or_rap_destroy.argtypes = [vp]
        L.or_rap_process.argtypes = [vp, C.POINTER(fp), fp, fp, C.c_int]
        L.or_rap_reset.argtypes = [vp]
        L.or_resample_output_count.argtypes = [C.c_int, C.c_double, C.c_double]
        L.or_resample_vgenp.argtypes = [fp, C.c_int, C.c_double, C.c_double, fp]
        L.or_biquad_make.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp]
        L.or_eq_prepare.restype = vp
        L.or_eq_prepare.argtypes = [C.c_double, dp, C.c_int, C.c_double, ip, ip, ip]
        L.or_eq_state_release.argtypes = [vp]
        L.or_eq_state_filter_count.argtypes = [vp]
        L.or_eq_state_coefficients.argtypes = [vp, C.c_int, dp]
        L.or_eq_state_preamp_linear.restype = C.c_double; L.or_eq_state_preamp_linear.argtypes = [vp]
        L.or_eq_state_reset.argtypes = [vp]
        L.or_eq_state_process.argtypes = [vp, fp, fp, fp, fp, C.c_int]
        L.or_eqp_create.restype = vp; L.or_eqp_create.argtypes = [C.c_double, C.c_int, ip]
        L.or_eqp_destroy.argtypes = [vp]
        L.or_eqp_set_target.argtypes = [vp, C.c_double, dp, C.c_int, ip, ip]
        L.or_eqp_reset.argtypes = [vp]
        L.or_eqp_drain_retired_states.argtypes = [vp]
        L.or_eqp_hold_publication_lock.argtypes = [vp, C.c_int]
        L.or_eqp_process.argtypes = [vp, fp, fp, fp, fp, C.c_int]
        L.or_direct_conv_f64.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, dp]
        L.or_synth_sample.restype = C.c_float
        L.or_synth_sample.argtypes = [C.c_uint32] * 4
        L.or_synth_fill.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, fp]
        L.or_bench_render.restype = C.c_double
        L.or_bench_render.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_uint32, dp]
        L.or_max_threads.restype = C.c_int
        L.or_batch_set_eq.argtypes = [vp, C.c_double, dp, C.c_int, C.c_double]
        L.or_batch_create.restype = vp
        L.or_batch_create.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_uint32]
        L.or_batch_step.restype = C.c_double
        L.or_batch_step.argtypes = [vp, C.c_int, C.c_int, dp]
        L.or_batch_destroy.argtypes = [vp]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class FFTSetup:
    """vDSP FFTSetup stand-in (vDSP_create_fftsetup; FFTSetupManager.swift:41-59)."""

    def __init__(self, log2n: int):
        self.log2n = log2n
        self._h = lib().or_fft_create(log2n)
        if not self._h:
            raise ValueError("invalid log2n")

    def zrip(self, re: np.ndarray, im: np.ndarray, forward: bool) -> None:
        assert re.dtype == np.float32 and im.dtype == np.float32
        lib().or_fft_zrip(self._h, _fp(re), _fp(im), 1 if forward else -1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_fft_destroy(self._h)
            self._h = None


class ConvolutionEngine:
    """ConvolutionEngine.swift:14-408."""

    def __init__(self, hrirSamples, blockSize: int = 512):
        self.hrirSamples = _f32(hrirSamples)
        self.blockSize = blockSize
        self._h = lib().or_conv_create(_fp(self.hrirSamples), len(self.hrirSamples), blockSize, None)
        if not self._h:
            raise ValueError("ConvolutionEngine init failed (init? returned nil)")

    @property
    def partitionCount(self) -> int:
        return lib().or_conv_partition_count(self._h)

    def process(self, input, frameCount: int | None = None) -> np.ndarray | None:
        """process(input:[Float], output:&, frameCount:) (:370-380): None when frameCount != blockSize."""
        x = _f32(input)
        out = np.zeros(self.blockSize, np.float32)
        count = self.blockSize if frameCount is None else frameCount
        if not lib().or_conv_process_array(self._h, _fp(x), _fp(out), count):
            return None
        return out

    def processAndAccumulate(self, input, outputAccumulator: np.ndarray) -> None:
        x = _f32(input)
        lib().or_conv_process_and_accumulate(self._h, _fp(x), _fp(outputAccumulator))

    def reset(self) -> None:
        lib().or_conv_reset(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_conv_destroy(self._h)
            self._h = None


class VirtualSpeakerRenderer:
    """HRIRManager.swift:84-88."""

    def __init__(self, speaker, convolverLeftEar: ConvolutionEngine, convolverRightEar: ConvolutionEngine):
        self.speaker = speaker
        self.convolverLeftEar = convolverLeftEar
        self.convolverRightEar = convolverRightEar


class RealtimeAudioProcessor:
    """RealtimeAudioProcessor.swift:11-191.

    ``literalStereo=True`` is the reference as written (min(renderers.count, 2), :145);
    ``False`` is the S-channel generalisation of SURVEY.md Q1 (renderer i <- input channel i).
    """

    def __init__(self, renderers, blockSize: int = 512, maxFramesPerCallback: int = 4096,
                 literalStereo: bool = True):
        self.renderers = list(renderers)
        self.blockSize = blockSize
        self.maxFramesPerCallback = maxFramesPerCallback
        self.literalStereo = literalStereo
        n = len(self.renderers)
        arr_t = C.c_void_p * max(n, 1)
        left = arr_t(*[r.convolverLeftEar._h for r in self.renderers])
        right = arr_t(*[r.convolverRightEar._h for r in self.renderers])
        self._h = lib().or_rap_create(left, right, n, blockSize, maxFramesPerCallback, 1 if literalStereo else 0)
        if not self._h:
            raise ValueError("precondition failed")
        self.inputCount = 2 if literalStereo else max(n, 1)

    def process(self, inputLeft, inputRight=None, frameCount: int | None = None):
        """Stereo entry (literal reference signature). Returns (left, right)."""
        return self.process_channels([inputLeft, inputRight], frameCount)

    def process_channels(self, inputs, frameCount: int | None = None, aliasOutputs: bool = False):
        arrs = [None if a is None else _f32(a) for a in inputs]
        n = len(arrs[0]) if frameCount is None else frameCount
        if n > self.maxFramesPerCallback:
            raise AssertionError("precondition(frameCount <= maxFramesPerCallback)")
        ptr_t = C.POINTER(C.c_float) * self.inputCount
        ptrs = ptr_t(*[(_fp(a) if a is not None else None) for a in arrs[: self.inputCount]])
        outL = np.full(n, np.nan, np.float32)
        outR = outL if aliasOutputs else np.full(n, np.nan, np.float32)
        lib().or_rap_process(self._h, ptrs, _fp(outL), _fp(outR), n)
        return outL, outR

    def reset(self) -> None:
        lib().or_rap_reset(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_rap_destroy(self._h)
            self._h = None


def resample_output_count(count: int, fromRate: float, toRate: float) -> int:
    return lib().or_resample_output_count(count, fromRate, toRate)


def resample_high_quality(input, fromRate: float, toRate: float) -> np.ndarray:
    """Resampler.resampleHighQuality (Resampler.swift:31-68) — vgenp semantics, parity unpinned."""
    x = _f32(input)
    if abs(fromRate - toRate) < 0.01:
        return x
    n = resample_output_count(len(x), fromRate, toRate)
    out = np.zeros(max(n, 0), np.float32)
    rc = lib().or_resample_vgenp(_fp(x), len(x), fromRate, toRate, _fp(out))
    if rc == -2:
        raise ValueError("down-sampling reads past the control vector in the reference (undefined)")
    return out[: max(rc, 0)]


def resample_linear_f64(input, fromRate: float, toRate: float) -> np.ndarray:
    """Checker for the flagged AW_RESAMPLE_CORRECT mode (no reference counterpart: Resampler.swift:16-30 documents linear
    interpolation at the target rate, its vgenp call computes something else — SURVEY.md Q7).  out[n] = input interpolated
    linearly at source position n * fromRate / toRate, float64 position and blend, last sample held, output count as
    Resampler.swift:39."""
    x = _f32(input)
    if abs(fromRate - toRate) < 0.01:
        return x
    n_out = resample_output_count(len(x), fromRate, toRate)
    if n_out <= 0:
        return np.zeros(0, np.float32)
    step = np.float64(fromRate) / np.float64(toRate)
    pos = np.arange(n_out, dtype=np.float64) * step
    m = np.minimum(pos.astype(np.int64), len(x) - 2) if len(x) > 1 else np.zeros(n_out, np.int64)
    xd = x.astype(np.float64)
    if len(x) == 1:
        return np.full(n_out, x[0], np.float32)
    frac = pos - m.astype(np.float64)
    y = (xd[m] + (xd[m + 1] - xd[m]) * frac).astype(np.float32)
    y[pos >= len(x) - 1] = x[-1]
    return y


class BiquadCoefficientError(Exception):
    NAMES = {1: "invalidSampleRate", 2: "invalidFrequency", 3: "invalidQ", 4: "nonFiniteInput",
             5: "nonFiniteCoefficients"}

    def __init__(self, code: int):
        super().__init__(self.NAMES.get(code, str(code)))
        self.code = code
        self.name = self.NAMES.get(code, str(code))


FILTER_TYPES = {"peaking": 0, "lowShelf": 1, "highShelf": 2, "PK": 0, "LSC": 1, "HSC": 2, 0: 0, 1: 1, 2: 2}


def biquad_make(type, gainDB: float, frequencyHz: float, q: float, sampleRate: float) -> np.ndarray:
    """BiquadCoefficientBuilder.make (BiquadCoefficientBuilder.swift:30-107) -> [b0,b1,b2,a1,a2]."""
    out = np.zeros(5, np.float64)
    rc = lib().or_biquad_make(FILTER_TYPES[type], gainDB, frequencyHz, q, sampleRate, _dp(out))
    if rc:
        raise BiquadCoefficientError(rc)
    return out


class ParametricEqualizerPreparationError(Exception):
    NAMES = {1: "invalidSampleRate", 2: "nonFinitePreamp", 3: "tooManyFilters", 4: "invalidFilter"}

    def __init__(self, code: int, index: int = -1, filter_error: int = 0):
        super().__init__(f"{self.NAMES.get(code, code)} index={index} error={filter_error}")
        self.code, self.index, self.filter_error = code, index, filter_error
        self.name = self.NAMES.get(code, str(code))


def _pack_filters(definition):
    """definition: None or dict(preampDB=float, filters=[dict(type, frequencyHz, gainDB, q, isEnabled)])."""
    if definition is None:
        return 0.0, np.zeros(0, np.float64), 0
    filters = definition.get("filters", [])
    arr = np.zeros((len(filters), 5), np.float64)
    for i, f in enumerate(filters):
        arr[i] = [FILTER_TYPES[f["type"]], 1.0 if f.get("isEnabled", True) else 0.0,
                  f["frequencyHz"], f["gainDB"], f["q"]]
    return float(definition.get("preampDB", 0.0)), arr.reshape(-1), len(filters)


class ParametricEqualizerState:
    """ParametricEqualizerProcessor.swift:16-98 (built through prepare, :174-217)."""

    def __init__(self, definition, sampleRate: float):
        preamp, arr, n = _pack_filters(definition)
        err, ei, ec = C.c_int(), C.c_int(), C.c_int()
        self._h = lib().or_eq_prepare(preamp, _dp(arr) if n else None, n, sampleRate,
                                      C.byref(err), C.byref(ei), C.byref(ec))
        if not self._h:
            raise ParametricEqualizerPreparationError(err.value, ei.value, ec.value)
        self.sampleRate = sampleRate

    @property
    def filterCount(self) -> int:
        return lib().or_eq_state_filter_count(self._h)

    @property
    def preampLinear(self) -> float:
        return lib().or_eq_state_preamp_linear(self._h)

    def coefficients(self, i: int) -> np.ndarray:
        out = np.zeros(5, np.float64)
        lib().or_eq_state_coefficients(self._h, i, _dp(out))
        return out

    def reset(self) -> None:
        lib().or_eq_state_reset(self._h)

    def process(self, left, right=None):
        l = _f32(left)
        r = None if right is None else _f32(right)
        outL = np.full(len(l), np.nan, np.float32)
        outR = np.full(len(l), np.nan, np.float32)
        lib().or_eq_state_process(self._h, _fp(l), None if r is None else _fp(r), _fp(outL), _fp(outR), len(l))
        return outL, outR

    def process_inplace(self, left: np.ndarray, right: np.ndarray, frameCount: int) -> None:
        lib().or_eq_state_process(self._h, _fp(left), _fp(right), _fp(left), _fp(right), frameCount)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_eq_state_release(self._h)
            self._h = None


class ParametricEqualizerProcessor:
    """ParametricEqualizerProcessor.swift:121-408."""

    @staticmethod
    def prepare(definition, sampleRate: float) -> ParametricEqualizerState:
        return ParametricEqualizerState(definition, sampleRate)

    def __init__(self, sampleRate: float, maxFramesPerCallback: int = 4096):
        err = C.c_int()
        self._h = lib().or_eqp_create(sampleRate, maxFramesPerCallback, C.byref(err))
        if not self._h:
            raise ParametricEqualizerPreparationError(err.value)
        self.sampleRate = sampleRate
        self.maxFramesPerCallback = maxFramesPerCallback

    def setTarget(self, definition) -> None:
        preamp, arr, n = _pack_filters(definition)
        ei, ec = C.c_int(), C.c_int()
        rc = lib().or_eqp_set_target(self._h, preamp, _dp(arr) if n else None, n, C.byref(ei), C.byref(ec))
        if rc:
            raise ParametricEqualizerPreparationError(rc, ei.value, ec.value)

    def reset(self) -> None:
        lib().or_eqp_reset(self._h)

    def drainRetiredStates(self) -> None:
        lib().or_eqp_drain_retired_states(self._h)

    def holdPublicationLock(self, held: bool) -> None:
        """withPublicationLockForTesting (:229-234) modelled as a flag."""
        lib().or_eqp_hold_publication_lock(self._h, 1 if held else 0)

    def process(self, left, right=None):
        l = _f32(left)
        r = None if right is None else _f32(right)
        if len(l) > self.maxFramesPerCallback:
            raise AssertionError("precondition(frameCount <= maxFramesPerCallback)")
        outL = np.full(len(l), np.nan, np.float32)
        outR = np.full(len(l), np.nan, np.float32)
        lib().or_eqp_process(self._h, _fp(l), None if r is None else _fp(r), _fp(outL), _fp(outR), len(l))
        return outL, outR

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_eqp_destroy(self._h)
            self._h = None


def direct_conv_f64(x, h) -> np.ndarray:
    """float64 direct convolution. x: [S][frames] f32; h: [S][2][taps] f32 -> out [2][frames] f64."""
    x = _f32(x)
    h = _f32(h)
    S, frames = x.shape
    assert h.shape[0] == S and h.shape[1] == 2
    out = np.zeros((2, frames), np.float64)
    lib().or_direct_conv_f64(_fp(x), S, frames, _fp(h), h.shape[2], _dp(out))
    return out


def synth_fill(seed: int, stream: int, speaker: int, frame0: int, frames: int) -> np.ndarray:
    out = np.zeros(frames, np.float32)
    lib().or_synth_fill(seed, stream, speaker, frame0, frames, _fp(out))
    return out


def synth_block(seed: int, streams, speakers: int, frame0: int, frames: int) -> np.ndarray:
    """[len(streams)][speakers][frames] synthetic input."""
    streams = list(streams)
    out = np.zeros((len(streams), speakers, frames), np.float32)
    for i, t in enumerate(streams):
        for s in range(speakers):
            lib().or_synth_fill(seed, t, s, frame0, frames, _fp(out[i, s]))
    return out


def bench_render(n_streams: int, S: int, B: int, h, blocks: int, threads: int = 0, seed: int = 0x41495257):
    """Time the reference-structured CPU render. h: [S][2][taps]. Returns (seconds, checksum)."""
    h = _f32(h)
    chk = C.c_double()
    sec = lib().or_bench_render(n_streams, S, B, _fp(h), h.shape[2], blocks, threads, seed, C.byref(chk))
    return sec, chk.value


def max_threads() -> int:
    return lib().or_max_threads()


class CpuBatch:
    """Persistent reference-structured CPU render of a bounded sample of streams (bench.py --impl reference)."""

    def __init__(self, n_streams: int, S: int, B: int, h, ring_blocks: int = 8, seed: int = 0x41495257):
        h = _f32(h)
        self.n_streams, self.S, self.B = n_streams, S, B
        self._h = lib().or_batch_create(n_streams, S, B, _fp(h), h.shape[2], ring_blocks, seed)

    def set_eq(self, definition, sampleRate: float) -> None:
        """Every stream of the sample gets its own ParametricEqualizerState after the spatial stage (full chain, C4)."""
        preamp, arr, n = _pack_filters(definition)
        rc = lib().or_batch_set_eq(self._h, preamp, _dp(arr) if n else None, n, sampleRate)
        if rc:
            raise ValueError(f"equalizer definition rejected ({rc})")

    def step(self, blocks: int = 1, threads: int = 0):
        chk = C.c_double()
        sec = lib().or_batch_step(self._h, blocks, threads, C.byref(chk))
        return sec, chk.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().or_batch_destroy(self._h)
            self._h = None
