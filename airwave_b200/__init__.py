"""airwave_b200 — B200-native batched binaural renderer behind Airwave's ConvolutionEngine /
RealtimeAudioProcessor / AudioEffectGraph interface.

The product is ``lib/libairwave_cuda.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/airwave_cuda.h``); this package is the thin host-side mirror of the reference's
interface used by tests and ``bench.py``.  There is no CPU fallback: importing the API without
the built library raises ImportError, and every ``*_create`` fails without a CUDA device.
"""
from ._lib import LIB_PATH, AirwaveError, declared_symbols, lib  # noqa: F401
from .api import *  # noqa: F401,F403

__version__ = "0.1.0"
