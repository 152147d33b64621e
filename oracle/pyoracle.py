"""ctypes front-end for the CPU oracle (oracle/oddio_oracle.hpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (oddio_b200/) never imports this.

Class and method names follow the reference (Frames, FramesSignal, Gain, Speed, Mixer,
SpatialScene, run, ...) so that the known-answer tests read like the reference's unit tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboddio_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("oddio_oracle_c.cpp", "oddio_oracle.hpp", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, f32, f64, u32, sz, i32 = C.c_void_p, C.c_float, C.c_double, C.c_uint32, C.c_size_t, C.c_int
    fp = C.POINTER(C.c_float)
    dp = C.POINTER(C.c_double)
    sig = {
        "orc_release": (None, [vp]),
        "orc_frames_new": (vp, [u32, i32, fp, sz]),
        "orc_frames_signal_new": (vp, [vp, f64]),
        "orc_cycle_new": (vp, [vp]),
        "orc_constant_new": (vp, [i32, fp]),
        "orc_sine_new": (vp, [f32, f32]),
        "orc_time_signal_new": (vp, [f32]),
        "orc_counting_signal_new": (vp, [u32]),
        "orc_finished_signal_new": (vp, []),
        "orc_mono_to_stereo_new": (vp, [vp]),
        "orc_speed_new": (vp, [vp]),
        "orc_speed_set": (None, [vp, f32]),
        "orc_speed_get": (f32, [vp]),
        "orc_gain_new": (vp, [vp]),
        "orc_gain_set_initial_ratio": (None, [vp, f32]),
        "orc_gain_set_initial_db": (None, [vp, f32]),
        "orc_gain_control_set_ratio": (None, [vp, f32]),
        "orc_gain_control_set_db": (None, [vp, f32]),
        "orc_gain_control_ratio": (f32, [vp]),
        "orc_gain_control_db": (f32, [vp]),
        "orc_fixed_gain_new": (vp, [vp, f32]),
        "orc_fixed_gain_value": (f32, [vp]),
        "orc_tanh_new": (vp, [vp]),
        "orc_reinhard_new": (vp, [vp]),
        "orc_signal_channels": (i32, [vp]),
        "orc_signal_sample": (None, [vp, f32, fp, sz]),
        "orc_run": (None, [vp, u32, fp, sz]),
        "orc_signal_is_finished": (i32, [vp]),
        "orc_signal_seek": (None, [vp, f32]),
        "orc_time_run": (f64, [vp, u32, fp, sz, i32, i32]),
        "orc_time_run_sharded": (f64, [C.POINTER(vp), i32, u32, fp, sz, i32, i32, i32]),
        "orc_frames_signal_t": (f64, [vp]),
        "orc_frames_signal_sample_t": (C.c_long, [vp]),
        "orc_frames_signal_playback_position": (f64, [vp]),
        "orc_frames_signal_control_is_finished": (i32, [vp]),
        "orc_cycle_cursor": (f64, [vp]),
        "orc_mixer_new": (vp, [i32]),
        "orc_mixer_play": (vp, [vp, vp]),
        "orc_mixed_stop": (None, [vp]),
        "orc_mixed_is_stopped": (i32, [vp]),
        "orc_mixer_len": (sz, [vp]),
        "orc_scene_new": (vp, []),
        "orc_scene_play": (vp, [vp, vp, fp, fp, f32]),
        "orc_scene_play_buffered": (vp, [vp, vp, fp, fp, f32, f32, u32, f32]),
        "orc_scene_set_listener_rotation": (None, [vp, fp]),
        "orc_scene_len": (sz, [vp, i32]),
        "orc_spatial_set_motion": (None, [vp, fp, fp, i32]),
        "orc_spatial_is_finished": (i32, [vp]),
        "orc_spatial_state": (None, [vp, fp]),
        "orc_out64": (sz, [vp, dp, sz]),
        "orc_ring_new": (vp, [sz]),
        "orc_ring_write": (None, [vp, vp, u32, f32]),
        "orc_ring_delay": (None, [vp, u32, f32]),
        "orc_ring_sample": (None, [vp, u32, f32, f32, fp, sz]),
        "orc_ring_write_cursor": (f32, [vp]),
        "orc_ring_buffer": (sz, [vp, fp, sz]),
        "orc_smoothed_new": (vp, [f32]),
        "orc_smoothed_set": (None, [vp, f32]),
        "orc_smoothed_advance": (None, [vp, f32]),
        "orc_smoothed_get": (f32, [vp]),
        "orc_smoothed_progress": (f32, [vp]),
        "orc_rotate": (None, [fp, fp, fp]),
        "orc_ear_state": (None, [fp, i32, f32, fp]),
        "orc_frames_interpolate": (f32, [vp, f64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return _fptr(np.ascontiguousarray(v, dtype=np.float32).reshape(-1))


class _Obj:
    def __init__(self, h):
        assert h, "oracle returned NULL"
        self._h = h

    def __del__(self):
        try:
            lib().orc_release(self._h)
        except Exception:
            pass


class Frames(_Obj):
    """frames.rs:19-22"""

    @staticmethod
    def from_slice(rate: int, samples) -> "Frames":
        a = np.ascontiguousarray(samples, dtype=np.float32)
        ch = 1 if a.ndim == 1 else a.shape[1]
        f = Frames(lib().orc_frames_new(int(rate), ch, _fptr(a), a.shape[0]))
        f.rate, f.channels, f.len = int(rate), ch, a.shape[0]
        return f

    def interpolate(self, s: float) -> float:
        return lib().orc_frames_interpolate(self._h, float(s))


class Signal(_Obj):
    """signal.rs:14-28 (+ Seek, :48-51)"""

    _keep = ()

    @property
    def channels(self) -> int:
        return lib().orc_signal_channels(self._h)

    def sample(self, interval: float, n: int) -> np.ndarray:
        ch = self.channels
        out = np.zeros((n, ch) if ch > 1 else (n,), dtype=np.float32)
        lib().orc_signal_sample(self._h, float(np.float32(interval)), _fptr(out), n)
        return out

    def sample_into(self, interval: float, out: np.ndarray) -> None:
        n = out.shape[0]
        lib().orc_signal_sample(self._h, float(np.float32(interval)), _fptr(out), n)

    def is_finished(self) -> bool:
        return bool(lib().orc_signal_is_finished(self._h))

    def seek(self, seconds: float) -> None:
        lib().orc_signal_seek(self._h, float(np.float32(seconds)))


def run(signal: Signal, sample_rate: int, n: int) -> np.ndarray:
    """lib.rs:90-93"""
    ch = signal.channels
    out = np.zeros((n, ch) if ch > 1 else (n,), dtype=np.float32)
    lib().orc_run(signal._h, int(sample_rate), _fptr(out), n)
    return out


def time_run(signal: Signal, sample_rate: int, n: int, warmup: int, reps: int) -> float:
    """Best-of-`reps` wall seconds of one run() call, timed inside C++ (steady_clock)."""
    out = np.zeros((n, signal.channels), dtype=np.float32)
    return lib().orc_time_run(signal._h, int(sample_rate), _fptr(out), n, warmup, reps)


def time_run_sharded(signals, sample_rate: int, n: int, warmup: int, reps: int):
    """Runs `signals` (aggregators holding disjoint shards of the sources) on one thread each and sums
    their tiles; returns (total wall seconds of `reps` rounds, last summed tile). Not reference behaviour."""
    ch = signals[0].channels
    out = np.zeros((n, ch), dtype=np.float32)
    arr = (C.c_void_p * len(signals))(*[s._h for s in signals])
    secs = lib().orc_time_run_sharded(arr, len(signals), int(sample_rate), _fptr(out), n, ch, warmup, reps)
    return secs, out


class FramesSignal(Signal):
    def __init__(self, frames: Frames, start_seconds: float = 0.0):
        super().__init__(lib().orc_frames_signal_new(frames._h, float(start_seconds)))
        self._keep = (frames,)

    @property
    def t(self) -> float:
        return lib().orc_frames_signal_t(self._h)

    @property
    def sample_t(self) -> int:
        return lib().orc_frames_signal_sample_t(self._h)

    def playback_position(self) -> float:
        return lib().orc_frames_signal_playback_position(self._h)

    def control_is_finished(self) -> bool:
        return bool(lib().orc_frames_signal_control_is_finished(self._h))


class Cycle(Signal):
    def __init__(self, frames: Frames):
        super().__init__(lib().orc_cycle_new(frames._h))
        self._keep = (frames,)

    @property
    def cursor(self) -> float:
        return lib().orc_cycle_cursor(self._h)


class Constant(Signal):
    def __init__(self, frame):
        a = np.atleast_1d(np.asarray(frame, dtype=np.float32))
        super().__init__(lib().orc_constant_new(a.shape[0], _fptr(a)))


class Sine(Signal):
    def __init__(self, phase: float, frequency_hz: float):
        super().__init__(lib().orc_sine_new(float(np.float32(phase)), float(np.float32(frequency_hz))))


class TimeSignal(Signal):  # ring.rs:86-97 test fixture
    def __init__(self, t0: float):
        super().__init__(lib().orc_time_signal_new(float(t0)))


class CountingSignal(Signal):  # signal.rs:97-108 test fixture
    def __init__(self, c0: int = 0):
        super().__init__(lib().orc_counting_signal_new(int(c0)))


class FinishedSignal(Signal):  # spatial.rs:611-627 test fixture
    def __init__(self):
        super().__init__(lib().orc_finished_signal_new())


class _Wrap(Signal):
    _ctor = None

    def __init__(self, inner: Signal, *args):
        super().__init__(getattr(lib(), self._ctor)(inner._h, *args))
        self._keep = (inner,)
        self.inner = inner


class MonoToStereo(_Wrap):
    _ctor = "orc_mono_to_stereo_new"


class Tanh(_Wrap):
    _ctor = "orc_tanh_new"


class Reinhard(_Wrap):
    _ctor = "orc_reinhard_new"


class Speed(_Wrap):
    _ctor = "orc_speed_new"

    def set_speed(self, factor: float) -> None:  # SpeedControl::set_speed
        lib().orc_speed_set(self._h, float(np.float32(factor)))

    def speed(self) -> float:
        return lib().orc_speed_get(self._h)


class FixedGain(_Wrap):
    _ctor = "orc_fixed_gain_new"

    def __init__(self, inner: Signal, db: float):
        super().__init__(inner, float(np.float32(db)))

    @property
    def gain(self) -> float:
        return lib().orc_fixed_gain_value(self._h)


class Gain(_Wrap):
    _ctor = "orc_gain_new"

    # Gain::set_gain / set_amplitude_ratio (initial, no smoothing) gain.rs:81-93
    def set_gain(self, db: float) -> None:
        lib().orc_gain_set_initial_db(self._h, float(np.float32(db)))

    def set_amplitude_ratio(self, factor: float) -> None:
        lib().orc_gain_set_initial_ratio(self._h, float(np.float32(factor)))

    # GainControl gain.rs:130-160
    def control_set_gain(self, db: float) -> None:
        lib().orc_gain_control_set_db(self._h, float(np.float32(db)))

    def control_set_amplitude_ratio(self, factor: float) -> None:
        lib().orc_gain_control_set_ratio(self._h, float(np.float32(factor)))

    def control_amplitude_ratio(self) -> float:
        return lib().orc_gain_control_ratio(self._h)

    def control_gain(self) -> float:
        return lib().orc_gain_control_db(self._h)


class Mixed(_Obj):
    """mixer.rs:30-44"""

    def stop(self) -> None:
        lib().orc_mixed_stop(self._h)

    def is_stopped(self) -> bool:
        return bool(lib().orc_mixed_is_stopped(self._h))


class Mixer(Signal):
    """mixer.rs:61-120; play() is MixerControl::play"""

    def __init__(self, channels: int = 2):
        super().__init__(lib().orc_mixer_new(channels))
        self._played = []

    def play(self, signal: Signal) -> Mixed:
        self._played.append(signal)
        return Mixed(lib().orc_mixer_play(self._h, signal._h))

    def __len__(self) -> int:
        return lib().orc_mixer_len(self._h)

    def out64(self, n: int) -> np.ndarray:
        ch = self.channels
        out = np.zeros(n * ch, dtype=np.float64)
        lib().orc_out64(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n * ch)
        return out.reshape(n, ch) if ch > 1 else out


class Spatial(_Obj):
    """spatial.rs:120-157"""

    def set_motion(self, position, velocity, discontinuity: bool) -> None:
        lib().orc_spatial_set_motion(self._h, _f3(position), _f3(velocity), int(bool(discontinuity)))

    def is_finished(self) -> bool:
        return bool(lib().orc_spatial_is_finished(self._h))

    def state(self) -> dict:
        out = np.zeros(8, dtype=np.float32)
        lib().orc_spatial_state(self._h, _fptr(out))
        return {
            "prev_position": out[0:3].copy(),
            "dt": out[3],
            "finished_for": out[4],
            "has_finished_for": bool(out[5]),
            "stopped": bool(out[6]),
        }


class SpatialScene(Signal):
    """spatial.rs:160-471; play()/play_buffered()/set_listener_rotation() are the SpatialSceneControl methods"""

    def __init__(self):
        super().__init__(lib().orc_scene_new())
        self._played = []

    def play(self, signal: Signal, position=(0, 0, 0), velocity=(0, 0, 0), radius: float = 0.1) -> Spatial:
        self._played.append(signal)
        return Spatial(lib().orc_scene_play(self._h, signal._h, _f3(position), _f3(velocity), float(np.float32(radius))))

    def play_buffered(self, signal: Signal, position, velocity, radius: float, max_distance: float, rate: int,
                      buffer_duration: float) -> Spatial:
        self._played.append(signal)
        return Spatial(
            lib().orc_scene_play_buffered(self._h, signal._h, _f3(position), _f3(velocity), float(np.float32(radius)),
                                          float(np.float32(max_distance)), int(rate), float(np.float32(buffer_duration))))

    def set_listener_rotation(self, q_xyzs) -> None:
        lib().orc_scene_set_listener_rotation(self._h, _f3(q_xyzs))

    def len(self, buffered: bool = False) -> int:
        return lib().orc_scene_len(self._h, int(buffered))

    def out64(self, n: int) -> np.ndarray:
        out = np.zeros(n * 2, dtype=np.float64)
        lib().orc_out64(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n * 2)
        return out.reshape(n, 2)


class Ring(_Obj):
    """ring.rs:4-80"""

    def __init__(self, capacity: int):
        super().__init__(lib().orc_ring_new(capacity))
        self.capacity = capacity

    def write(self, signal: Signal, rate: int, dt: float) -> None:
        lib().orc_ring_write(self._h, signal._h, int(rate), float(np.float32(dt)))

    def delay(self, rate: int, dt: float) -> None:
        lib().orc_ring_delay(self._h, int(rate), float(np.float32(dt)))

    def sample(self, rate: int, t: float, interval: float, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float32)
        lib().orc_ring_sample(self._h, int(rate), float(np.float32(t)), float(np.float32(interval)), _fptr(out), n)
        return out

    @property
    def write_cursor(self) -> float:
        return lib().orc_ring_write_cursor(self._h)

    @property
    def buffer(self) -> np.ndarray:
        out = np.zeros(self.capacity, dtype=np.float32)
        lib().orc_ring_buffer(self._h, _fptr(out), self.capacity)
        return out


class Smoothed(_Obj):
    """smooth.rs:26-72"""

    def __init__(self, x: float):
        super().__init__(lib().orc_smoothed_new(float(np.float32(x))))

    def set(self, v: float) -> None:
        lib().orc_smoothed_set(self._h, float(np.float32(v)))

    def advance(self, p: float) -> None:
        lib().orc_smoothed_advance(self._h, float(np.float32(p)))

    def get(self) -> float:
        return lib().orc_smoothed_get(self._h)

    def progress(self) -> float:
        return lib().orc_smoothed_progress(self._h)


def rotate(q_xyzs, p) -> np.ndarray:
    """math/mod.rs:81-94"""
    out = np.zeros(3, dtype=np.float32)
    lib().orc_rotate(_f3(q_xyzs), _f3(p), _fptr(out))
    return out


def ear_state(p, ear: int, radius: float):
    """spatial.rs:531-549 -> (offset, gain)"""
    out = np.zeros(2, dtype=np.float32)
    lib().orc_ear_state(_f3(p), int(ear), float(np.float32(radius)), _fptr(out))
    return out[0], out[1]
