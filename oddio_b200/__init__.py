"""oddio_b200 — B200-native implementation of Ralith/oddio's spatial/mixer hot path.

The product is `liboddio_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/oddio_b200.h). This package is the host-side mirror of the reference's interface over
that ABI (see api.py). There is no CPU implementation here; the CPU oracle lives under oracle/
and is test infrastructure only.
"""
from ._lib import OddioError, SO_PATH, load  # noqa: F401
from .api import (  # noqa: F401
    Context, Cycle, FixedGain, Frames, FramesSignal, FramesSignalControl, Gain, GainControl, Mixed, Mixer, MixerControl,
    Reinhard, Signal, Spatial, SpatialOptions, SpatialScene, SpatialSceneControl, Speed, SpeedControl, Tanh,
    default_context, flatten_stereo, frame_stereo, init, run,
)

from . import wavio  # noqa: F401,E402

Sample = "f32"  # lib.rs:85
