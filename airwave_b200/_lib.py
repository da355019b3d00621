"""ctypes binding of libairwave_cuda.so (include/airwave_cuda.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is missing
or a symbol the header declares is absent, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# AW_LIBRARY: developer override to load an experimental build of the same library (kernel A/B tests); never a fallback
LIB_PATH = os.environ.get("AW_LIBRARY") or os.path.join(HERE, "lib", "libairwave_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "airwave_cuda.h")


class AirwaveError(RuntimeError):
    """Raised for any non-zero aw_status; carries the status code and aw_last_error()."""

    def __init__(self, status: int, message: str, bad_index: int = -1, bad_reason: int = 0):
        super().__init__(f"[aw_status {status}] {message}")
        self.status = status
        self.message = message
        self.bad_index = bad_index
        self.bad_reason = bad_reason


class EngineConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_streams", C.c_int32), ("n_speakers", C.c_int32), ("block", C.c_int32),
                ("sample_rate", C.c_double), ("max_frames_per_call", C.c_int32), ("max_partitions", C.c_int32),
                ("flags", C.c_uint32)]


class EqFilter(C.Structure):
    _fields_ = [("type", C.c_int32), ("enabled", C.c_int32), ("frequency_hz", C.c_double), ("gain_db", C.c_double),
                ("q", C.c_double), ("source_line", C.c_int32), ("source_number", C.c_int32)]


# status codes (aw_status)
OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_INVALID_BLOCK_SIZE, ERR_FRAME_COUNT = 1, 2, 3, 4, 5
ERR_CHANNEL_MAPPING, ERR_NO_RENDERERS, ERR_RANGE, ERR_MISMATCH, ERR_UNSUPPORTED, ERR_RESAMPLE_DOWN, ERR_NOT_READY = 6, 7, 8, 9, 10, 11, 12
ERR_EQ_INVALID_SAMPLE_RATE, ERR_EQ_NON_FINITE_PREAMP, ERR_EQ_TOO_MANY_FILTERS, ERR_EQ_INVALID_FILTER = 20, 21, 22, 23
ERR_WAV_READ, ERR_WAV_CHANNEL_COUNT, ERR_WAV_EMPTY, ERR_WAV_UNSUPPORTED_FORMAT, ERR_EQ_PARSE = 30, 31, 32, 33, 40
ENGINE_LITERAL_STEREO, ENGINE_PIPELINED, ENGINE_OVERLAP_EQ = 1, 2, 4
RESET_SPATIAL, RESET_EQ = 1, 2
RESAMPLE_REFERENCE, RESAMPLE_CORRECT = 0, 1
LAYOUT_STEREO, LAYOUT_SURROUND51, LAYOUT_SURROUND71, LAYOUT_ATMOS714 = 2, 6, 8, 12
FILTER_TYPES = {"peaking": 0, "lowShelf": 1, "highShelf": 2, "PK": 0, "LSC": 1, "HSC": 2, 0: 0, 1: 1, 2: 2}
FILTER_NAMES = {0: "peaking", 1: "lowShelf", 2: "highShelf"}


def declared_symbols() -> list[str]:
    """Every AW_API function include/airwave_cuda.h declares."""
    text = open(HEADER_PATH).read()
    return re.findall(r"AW_API\s+[^;(]*?\b(aw_[a-z0-9_]+)\s*\(", text)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m airwave_b200.build` "
                          "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise ImportError(f"libairwave_cuda.so lacks symbols declared in airwave_cuda.h: {missing}")
    fp, ip, dp, vp, ll = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p, C.c_longlong
    ull = C.POINTER(C.c_ulonglong)
    L.aw_version.restype = C.c_char_p
    L.aw_status_string.restype = C.c_char_p; L.aw_status_string.argtypes = [C.c_int]
    L.aw_last_error.restype = C.c_char_p
    L.aw_plan_prepare.argtypes = [C.c_int, C.c_int]
    L.aw_plan_cache_stats.argtypes = [C.c_int, ip, ip, C.c_int]
    L.aw_wav_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.aw_wav_load_memory.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.aw_wav_info.argtypes = [vp, dp, ip, ip]
    L.aw_wav_channel.restype = fp; L.aw_wav_channel.argtypes = [vp, C.c_int]
    L.aw_wav_destroy.argtypes = [vp]; L.aw_wav_destroy.restype = None
    L.aw_layout_speakers.argtypes = [C.c_int, ip, C.c_int]
    L.aw_hesuvi_map.argtypes = [C.c_int, ip, C.c_int, ip, ip]
    L.aw_hesuvi_parse.argtypes = [C.c_char_p, ip, ip]
    L.aw_resample_output_count.argtypes = [C.c_int, C.c_double, C.c_double]
    L.aw_resample.argtypes = [C.c_int, fp, C.c_int, C.c_double, C.c_double, fp, C.c_int, ip]
    L.aw_resample_ex.argtypes = [C.c_int, fp, C.c_int, C.c_double, C.c_double, C.c_int, fp, C.c_int, ip]
    L.aw_bank_create.argtypes = [C.c_int, fp, C.c_int, C.c_int, C.c_double, C.c_double, ip, ip, C.c_int, C.c_int, C.POINTER(vp)]
    L.aw_bank_create_ex.argtypes = [C.c_int, fp, C.c_int, C.c_int, C.c_double, C.c_double, ip, ip, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.aw_bank_create_from_wav.argtypes = [C.c_int, vp, C.c_double, C.c_int, C.c_int, C.POINTER(vp)]
    L.aw_bank_info.argtypes = [vp, ip, ip, ip, ip]
    L.aw_bank_rows.argtypes = [vp]
    L.aw_bank_read.argtypes = [vp, fp, fp]
    L.aw_bank_destroy.argtypes = [vp]; L.aw_bank_destroy.restype = None
    L.aw_engine_create.argtypes = [C.POINTER(EngineConfig), C.POINTER(vp)]
    L.aw_engine_destroy.argtypes = [vp]; L.aw_engine_destroy.restype = None
    L.aw_engine_set_bank.argtypes = [vp, C.c_int, C.c_int, vp]
    eq_args = [vp, C.c_int, C.c_int, C.c_double, C.POINTER(EqFilter), C.c_int]
    L.aw_engine_eq_prepare.argtypes = eq_args + [ip, ip]
    L.aw_engine_eq_update.argtypes = eq_args + [ip, ip]
    L.aw_engine_eq_set_target.argtypes = eq_args + [C.c_int, ip, ip]
    L.aw_engine_eq_install_state.argtypes = eq_args + [ip, ip]
    L.aw_engine_eq_drain_retired.argtypes = [vp, C.c_int, C.c_int]
    L.aw_engine_eq_active.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.aw_engine_eq_hold_publication.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.aw_engine_process.argtypes = [vp, vp, vp, C.c_int]
    L.aw_engine_process_device.argtypes = [vp, vp, ll, ll, vp, ll, ll, C.c_int]
    L.aw_engine_process_stereo.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    L.aw_engine_submit.argtypes = [vp, vp, vp, C.c_int]
    L.aw_engine_submit_device.argtypes = [vp, vp, ll, ll, vp, C.c_int]
    L.aw_engine_wait.argtypes = [vp]
    L.aw_engine_flush.argtypes = [vp]
    L.aw_engine_reset.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.aw_engine_counters.argtypes = [vp, ull, ull, ull, ull]
    L.aw_engine_stream.restype = vp; L.aw_engine_stream.argtypes = [vp]
    L.aw_engine_uses_tensor_maps.argtypes = [vp]
    L.aw_engine_profile_begin.argtypes = [vp, C.c_int]
    L.aw_engine_plan.argtypes = [vp, ip, ip, ip]
    L.aw_engine_kernels.argtypes = [vp, C.c_char_p, C.c_int]
    L.aw_engine_profile_end.argtypes = [vp, dp, ull]
    L.aw_host_alloc.restype = vp; L.aw_host_alloc.argtypes = [C.c_size_t]
    L.aw_host_free.argtypes = [vp]; L.aw_host_free.restype = None
    L.aw_biquad_make.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp]
    L.aw_eq_parse.argtypes = [C.c_char_p, C.c_size_t, dp, C.POINTER(EqFilter), C.c_int, ip, C.c_char_p, C.c_size_t]
    L.aw_synth_fill_device.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, ll, C.c_int, C.c_uint32, vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().aw_last_error().decode("utf-8", "replace")


def check(status: int, bad_index: int = -1, bad_reason: int = 0) -> None:
    if status != OK:
        raise AirwaveError(status, last_error() or lib().aw_status_string(status).decode(), bad_index, bad_reason)
