"""CPU oracle for the Airwave binaural render path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or as the timed
CPU baseline.  The product (``airwave_b200``) never imports it.

``airwave_oracle.c`` restates the reference's DSP (ConvolutionEngine, RealtimeAudioProcessor,
Resampler, BiquadCoefficientBuilder, ParametricEqualizerState/Processor) in plain C;
``host.py`` restates the setup-time host logic (WAVLoader, HRIRChannelMap, InputLayout,
EqualizerAPOParser, AudioEffectGraph routing, HRIRManager activation) in pure Python.
Every function cites the reference file:line it follows.
"""
from .binding import *  # noqa: F401,F403
from .host import *  # noqa: F401,F403
