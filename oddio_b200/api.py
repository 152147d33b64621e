"""Host-side mirror of the reference's interface for the spatial/mixer hot path, over the C ABI.

Names, argument meaning and control-flow semantics follow Ralith/oddio 0.7.4 (file:line cited per
item, all under the reference's src/). The reference is Rust and there is no rustc in this image,
so this Python layer plays the role the Rust shim crate of INTEGRATION.md plays for a Rust caller;
both sit on the same ``extern "C"`` entry points and contain no arithmetic of the hot path.

Differences forced by the device path (documented in DESIGN.md):
  * the set of playable signals is closed: FramesSignal, optionally under Speed / FixedGain / Gain
    (SURVEY.md §7 H3); anything else raises OddioError(ODB_E_UNSUPPORTED);
  * signals are descriptions until played; `play` moves them onto the device.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import Chain, OddioError, check

_default_ctx: Optional["Context"] = None


class Context:
    """One CUDA device + stream + PCM arena (odb_ctx). One per process and GPU."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        L = _lib.load()
        h = C.c_void_p()
        if stream is None:
            check(L.odb_ctx_create(int(device), C.byref(h)))
        else:
            check(L.odb_ctx_create_on_stream(int(device), C.c_void_p(stream), C.byref(h)))
        self._h = h
        self.device = int(device)

    def synchronize(self) -> None:
        check(_lib.load().odb_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        out = C.c_void_p()
        check(_lib.load().odb_ctx_stream(self._h, C.byref(out)))
        return out.value or 0

    def close(self) -> None:
        if self._h:
            _lib.load().odb_ctx_destroy(self._h)
            self._h = None


def init(device: Optional[int] = None, stream: Optional[int] = None) -> Context:
    """Creates (or returns) the process-wide default context. `device` defaults to LOCAL_RANK or 0."""
    global _default_ctx
    if _default_ctx is None:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        _default_ctx = Context(device, stream)
    return _default_ctx


def default_context() -> Context:
    return init()


def _f32(x) -> float:
    return float(np.float32(x))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _vec3(v) -> np.ndarray:
    a = np.ascontiguousarray(v, dtype=np.float32).reshape(-1)
    if a.shape[0] != 3:
        raise ValueError("expected a 3-vector")
    return a


# ---------------------------------------------------------------------------------------------
class Frames:
    """`Arc<Frames<T>>` (frames.rs:19-77): PCM resident in HBM. channels 1 = Sample, 2 = [Sample; 2]."""

    def __init__(self, ctx: Context, handle: int, rate: int, channels: int, n_frames: int):
        self._ctx, self._h = ctx, handle
        self._rate, self.channels, self._len = int(rate), int(channels), int(n_frames)

    @staticmethod
    def from_slice(rate: int, samples, ctx: Optional[Context] = None) -> "Frames":
        """Frames::from_slice (frames.rs:26-47). `samples`: shape (n,) mono or (n, 2) stereo, host memory."""
        ctx = ctx or default_context()
        a = np.ascontiguousarray(samples, dtype=np.float32)
        ch = 1 if a.ndim == 1 else int(a.shape[1])
        h = C.c_uint64()
        check(_lib.load().odb_frames_from_slice(ctx._h, int(rate), ch, _fptr(a), a.shape[0], C.byref(h)))
        return Frames(ctx, h.value, rate, ch, a.shape[0])

    @staticmethod
    def from_i16(rate: int, samples, bits_per_sample: int = 16, ctx: Optional[Context] = None) -> "Frames":
        """Integer PCM (examples/wav.rs:30-46): uploaded as int16 and scaled to f32 on the device,
        `sample as f32 / (2^(bits-1) - 1) as f32`. `samples`: int16, shape (n,) or (n, 2)."""
        ctx = ctx or default_context()
        a = np.ascontiguousarray(samples, dtype=np.int16)
        ch = 1 if a.ndim == 1 else int(a.shape[1])
        h = C.c_uint64()
        check(_lib.load().odb_frames_from_i16(ctx._h, int(rate), ch, a.ctypes.data_as(C.POINTER(C.c_int16)), a.shape[0],
                                              int(bits_per_sample), C.byref(h)))
        return Frames(ctx, h.value, rate, ch, a.shape[0])

    @staticmethod
    def from_iter(rate: int, it, ctx: Optional[Context] = None) -> "Frames":
        """Frames::from_iter (frames.rs:50-77)."""
        return Frames.from_slice(rate, np.asarray(list(it), dtype=np.float32), ctx)

    @staticmethod
    def from_device(rate: int, channels: int, dev_ptr: int, n_frames: int, ctx: Optional[Context] = None) -> "Frames":
        """PCM already in device memory (decoded / synthesised on the GPU); copied into the arena."""
        ctx = ctx or default_context()
        h = C.c_uint64()
        check(_lib.load().odb_frames_from_device(ctx._h, int(rate), int(channels), C.c_void_p(dev_ptr), int(n_frames), C.byref(h)))
        return Frames(ctx, h.value, rate, channels, n_frames)

    def rate(self) -> int:  # frames.rs:80-83
        return self._rate

    def __len__(self) -> int:
        return self._len

    def release(self) -> None:
        if self._h:
            _lib.load().odb_frames_release(self._ctx._h, self._h)
            self._h = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# The closed set of signal descriptions. Controls are bound when played.
class _Bound:
    """A control handle that starts working once its signal is playing (owner + source id)."""

    def __init__(self):
        self._owner = None  # odb_scene* / odb_mixer*
        self._src = 0

    def _bind(self, owner, src: int) -> None:
        self._owner, self._src = owner, src

    def _need(self):
        if self._owner is None:
            raise OddioError(_lib.ODB_E_INVALID, "signal is not playing yet")
        return self._owner, self._src


class Signal:
    """signal.rs:14-28. Descriptions have no `sample`; only the aggregators (SpatialScene, Mixer) do."""

    channels = 1

    def _chain(self, chain: Chain, controls: list) -> None:
        raise OddioError(_lib.ODB_E_UNSUPPORTED, f"{type(self).__name__} cannot run on the device path")

    _is_seek = False


class FramesSignalControl(_Bound):
    """frames.rs:229-248"""

    def __init__(self, frames: Frames, start_seconds: float):
        super().__init__()
        self._frames, self._start = frames, float(start_seconds)

    def playback_position(self) -> float:
        if self._owner is None:  # not playing yet: sample_t as initialised by FramesSignal::new (frames.rs:160)
            return float(int(self._start * self._frames.rate())) / self._frames.rate()
        out = C.c_double()
        check(_lib.load().odb_source_playback_position(self._owner, self._src, C.byref(out)))
        return out.value

    def is_finished(self) -> bool:
        if self._owner is None:
            st = int(self._start * self._frames.rate())
            return st >= 0 and st >= len(self._frames)
        out = C.c_int()
        check(_lib.load().odb_source_frames_is_finished(self._owner, self._src, C.byref(out)))
        return bool(out.value)

    def cursor(self) -> Tuple[float, float]:
        """Parity aid: (FramesSignal::t as f64, Ring::write as f32)."""
        owner, src = self._need()
        t, w = C.c_double(), C.c_float()
        check(_lib.load().odb_source_cursor(owner, src, C.byref(t), C.byref(w)))
        return t.value, w.value


class FramesSignal(Signal):
    """frames.rs:141-220"""

    _is_seek = True

    def __init__(self, frames: Frames, start_seconds: float = 0.0):
        self.frames, self.start_seconds = frames, float(start_seconds)
        self.channels = frames.channels
        self.control = FramesSignalControl(frames, start_seconds)

    @staticmethod
    def new(frames: Frames, start_seconds: float = 0.0) -> Tuple[FramesSignalControl, "FramesSignal"]:
        """FramesSignal::new (frames.rs:156-169) -> (control, signal)"""
        s = FramesSignal(frames, start_seconds)
        return s.control, s

    def _chain(self, chain: Chain, controls: list) -> None:
        chain.frames = self.frames._h
        chain.start_seconds = self.start_seconds
        controls.append(self.control)


class Cycle(Signal):
    """cycle.rs:6-61: loops a Frames block end to end. On the device it plays under a Mixer (optionally inside
    Speed / FixedGain / Gain); the literal kernel walks its cursor exactly as Cycle::sample does."""

    _is_seek = True

    def __init__(self, frames: Frames):
        self.frames, self.channels = frames, frames.channels
        self._cursor = 0.0  # in samples (cycle.rs:8); moved by seek() until the signal is played
        self.control = FramesSignalControl(frames, 0.0)  # parity aid only: cursor() reads the device's f64 cursor

    def seek(self, seconds: float) -> None:
        """Seek::seek (cycle.rs:57-60), before playing."""
        n = float(len(self.frames))
        c = self._cursor + float(np.float32(seconds)) * float(self.frames.rate())
        r = float(np.fmod(c, n))
        self._cursor = r + n if r < 0.0 else r  # f64::rem_euclid

    def _chain(self, chain: Chain, controls: list) -> None:
        chain.frames = self.frames._h
        chain.start_seconds = self._cursor
        chain.flags |= _lib.CHAIN_CYCLE
        controls.append(self.control)


class SpeedControl(_Bound):
    """speed.rs:43-55"""

    def __init__(self):
        super().__init__()
        self._speed = 1.0

    def speed(self) -> float:
        return self._speed

    def set_speed(self, factor: float) -> None:
        self._speed = _f32(factor)
        if self._owner is not None:
            check(_lib.load().odb_source_set_speed(self._owner, self._src, self._speed))


class Speed(Signal):
    """speed.rs:9-41. Not Seek."""

    def __init__(self, inner: Signal):
        if not isinstance(inner, (FramesSignal, Cycle)):
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "device path: Speed must wrap a FramesSignal or Cycle directly")
        self.inner, self.channels = inner, inner.channels
        self.control = SpeedControl()

    @staticmethod
    def new(inner: Signal) -> Tuple[SpeedControl, "Speed"]:
        s = Speed(inner)
        return s.control, s

    def _chain(self, chain: Chain, controls: list) -> None:
        self.inner._chain(chain, controls)
        chain.flags |= _lib.CHAIN_SPEED
        chain.speed = self.control._speed
        controls.append(self.control)


class FixedGain(Signal):
    """gain.rs:9-51. Seek iff inner is."""

    def __init__(self, inner: Signal, db: float):
        if not isinstance(inner, (FramesSignal, Cycle, Speed)):
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "device path: FixedGain must wrap FramesSignal or Speed")
        self.inner, self.channels, self.db = inner, inner.channels, _f32(db)
        self._is_seek = inner._is_seek

    def _chain(self, chain: Chain, controls: list) -> None:
        self.inner._chain(chain, controls)
        chain.flags |= _lib.CHAIN_FIXED_GAIN
        chain.fixed_gain_db = self.db


class GainControl(_Bound):
    """gain.rs:130-160"""

    def __init__(self):
        super().__init__()
        self._ratio = np.float32(1.0)

    def gain(self) -> float:
        return float(np.float32(20.0) * np.log10(self._ratio, dtype=np.float32))

    def set_gain(self, db: float) -> None:
        self.set_amplitude_ratio(float(np.power(np.float32(10.0), np.float32(db) / np.float32(20.0), dtype=np.float32)))

    def amplitude_ratio(self) -> float:
        return float(self._ratio)

    def set_amplitude_ratio(self, factor: float) -> None:
        self._ratio = np.float32(factor)
        if self._owner is not None:
            check(_lib.load().odb_source_set_amplitude_ratio(self._owner, self._src, float(self._ratio)))


class Gain(Signal):
    """gain.rs:58-127. Not Seek."""

    def __init__(self, inner: Signal):
        if not isinstance(inner, (FramesSignal, Cycle, Speed, FixedGain)):
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "device path: Gain must wrap FramesSignal, Speed or FixedGain")
        self.inner, self.channels = inner, inner.channels
        self.control = GainControl()

    @staticmethod
    def new(inner: Signal) -> Tuple[GainControl, "Gain"]:
        s = Gain(inner)
        return s.control, s

    def set_amplitude_ratio(self, factor: float) -> None:
        """Gain::set_amplitude_ratio (gain.rs:90-93): immediate, before playing."""
        self.control._ratio = np.float32(factor)

    def set_gain(self, db: float) -> None:
        """Gain::set_gain (gain.rs:81-83)"""
        self.control._ratio = np.power(np.float32(10.0), np.float32(db) / np.float32(20.0), dtype=np.float32)

    def _chain(self, chain: Chain, controls: list) -> None:
        self.inner._chain(chain, controls)
        chain.flags |= _lib.CHAIN_GAIN
        chain.gain_ratio = float(self.control._ratio)
        controls.append(self.control)


def _build_chain(signal: Signal):
    chain, controls = Chain(), []
    chain.speed, chain.gain_ratio = 1.0, 1.0
    signal._chain(chain, controls)
    return chain, controls


# ---------------------------------------------------------------------------------------------
@dataclass
class SpatialOptions:
    """spatial.rs:354-371"""

    position: Sequence[float] = (0.0, 0.0, 0.0)
    velocity: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 0.1


class Spatial:
    """spatial.rs:120-157"""

    def __init__(self, scene_h, src: int):
        self._scene, self._src = scene_h, src

    def set_motion(self, position, velocity, discontinuity: bool) -> None:
        p, v = _vec3(position), _vec3(velocity)
        check(_lib.load().odb_spatial_set_motion(self._scene, self._src, _fptr(p), _fptr(v), int(bool(discontinuity))))

    def is_finished(self) -> bool:
        out = C.c_int()
        check(_lib.load().odb_spatial_is_finished(self._scene, self._src, C.byref(out)))
        return bool(out.value)


class _Aggregator(Signal):
    """Common part of the two hot-loop owners: sample / run / epilogue wrapper."""

    _sample = _sample_device = _sample_i16 = _destroy = _set_epilogue = None

    def _out(self, n: int, out: Optional[np.ndarray]) -> np.ndarray:
        shape = (n, self.channels) if self.channels > 1 else (n,)
        if out is None:
            return np.zeros(shape, dtype=np.float32)
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != n * self.channels:
            raise ValueError("out must be a C-contiguous float32 array of n frames")
        return out

    def sample(self, interval: float, n_or_out) -> np.ndarray:
        """Signal::sample (signal.rs:19): `n_or_out` is a frame count or the buffer to fill."""
        if isinstance(n_or_out, np.ndarray):
            out = self._out(n_or_out.shape[0], n_or_out)
        else:
            out = self._out(int(n_or_out), None)
        check(self._sample(self._h, _f32(interval), _fptr(out), out.shape[0]))
        return out

    def sample_device(self, interval: float, dev_ptr: int, n_frames: int) -> None:
        check(self._sample_device(self._h, _f32(interval), C.c_void_p(dev_ptr), int(n_frames)))

    def sample_i16(self, interval: float, n_frames: int) -> np.ndarray:
        """One callback quantised to 16-bit PCM on the device, `(sample * i16::MAX as f32) as i16`
        (examples/offline.rs:39); returns int16 of shape (n, channels) or (n,)."""
        n = int(n_frames)
        out = np.zeros((n, self.channels) if self.channels > 1 else (n,), dtype=np.int16)
        check(self._sample_i16(self._h, _f32(interval), out.ctypes.data_as(C.POINTER(C.c_int16)), n))
        return out

    def is_finished(self) -> bool:  # spatial.rs:473-476 / Signal default
        return False

    def last_launch_count(self) -> int:
        out = C.c_uint32()
        check(_lib.load().odb_last_launch_count(self._h, C.byref(out)))
        return out.value

    def last_job_counters(self) -> dict:
        """(source, tile) jobs of the last callback by kernel: literal general kernel, staged / streaming kernel,
        staged resampling kernel (mixer only), literal ring kernel (buffered sources whose reads wrap)."""
        out = (C.c_uint32 * 4)()
        check(_lib.load().odb_last_job_counters(self._h, out))
        return {"general": int(out[0]), "staged": int(out[1]), "resampled": int(out[2]), "ring_literal": int(out[3])}

    def set_profiling(self, enabled: bool) -> None:
        check(_lib.load().odb_set_profiling(self._h, int(bool(enabled))))

    def last_mix_kernel_ms(self) -> float:
        out = C.c_float()
        check(_lib.load().odb_last_mix_kernel_ms(self._h, C.byref(out)))
        return out.value

    def set_kernel_variant(self, variant: int) -> None:
        check(_lib.load().odb_set_kernel_variant(self._h, int(variant)))

    def close(self) -> None:
        if self._h:
            self._destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpatialScene(_Aggregator):
    """spatial.rs:160-189, :373-477 — the audio-thread half."""

    channels = 2

    def sample_exchange(self, exchange, interval: float, dev_ptr: int, n_frames: int, lag: int = 0, epilogue: int = 0) -> bool:
        """One callback of this rank's shard with the multi-GPU exchange folded into the callback kernel
        (odb_scene_sample_exchange). Returns True when `dev_ptr` received a summed tile (callback k - lag)."""
        w = C.c_int(0)
        check(_lib.load().odb_scene_sample_exchange(self._h, exchange._h, _f32(interval), C.c_void_p(dev_ptr), int(n_frames),
                                                    int(lag), int(epilogue), C.byref(w)))
        return bool(w.value)

    def __init__(self, ctx: Context):
        L = _lib.load()
        h = C.c_void_p()
        check(L.odb_scene_create(ctx._h, C.byref(h)))
        self._h, self._ctx = h, ctx
        self._sample, self._sample_device, self._sample_i16 = L.odb_scene_sample, L.odb_scene_sample_device, L.odb_scene_sample_i16
        self._destroy, self._set_epilogue = L.odb_scene_destroy, L.odb_scene_set_epilogue

    @staticmethod
    def new(ctx: Optional[Context] = None) -> Tuple["SpatialSceneControl", "SpatialScene"]:
        """SpatialScene::new (spatial.rs:170-188) -> (control, signal)"""
        scene = SpatialScene(ctx or default_context())
        return SpatialSceneControl(scene), scene

    def len(self, buffered: bool = False) -> int:
        out = C.c_uint64()
        check(_lib.load().odb_scene_len(self._h, int(buffered), C.byref(out)))
        return out.value


class SpatialSceneControl:
    """spatial.rs:268-350 — the control-thread half."""

    def __init__(self, scene: SpatialScene):
        self._scene = scene

    def play(self, signal: Signal, options: SpatialOptions) -> Spatial:
        """spatial.rs:289-302; requires Seek (FramesSignal, optionally under FixedGain)."""
        if signal.channels != 1:
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "spatial signals are mono (Frame = Sample)")
        chain, controls = _build_chain(signal)
        p, v = _vec3(options.position), _vec3(options.velocity)
        src = C.c_uint64()
        check(_lib.load().odb_scene_play(self._scene._h, C.byref(chain), _fptr(p), _fptr(v), _f32(options.radius), C.byref(src)))
        for c in controls:
            c._bind(self._scene._h, src.value)
        return Spatial(self._scene._h, src.value)

    def play_buffered(self, signal: Signal, options: SpatialOptions, max_distance: float, rate: int,
                      buffer_duration: float) -> Spatial:
        """spatial.rs:314-340"""
        if signal.channels != 1:
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "spatial signals are mono (Frame = Sample)")
        chain, controls = _build_chain(signal)
        p, v = _vec3(options.position), _vec3(options.velocity)
        src = C.c_uint64()
        check(_lib.load().odb_scene_play_buffered(self._scene._h, C.byref(chain), _fptr(p), _fptr(v), _f32(options.radius),
                                                  _f32(max_distance), int(rate), _f32(buffer_duration), C.byref(src)))
        for c in controls:
            c._bind(self._scene._h, src.value)
        return Spatial(self._scene._h, src.value)

    def set_motion_many(self, spatials: Sequence["Spatial"], positions, velocities, discontinuity=None) -> None:
        """Spatial::set_motion for many sources in one foreign call (odb_spatial_set_motion_many)."""
        n = len(spatials)
        ids = (C.c_uint64 * n)(*[s._src for s in spatials])
        self.set_motion_ids(ids, n, positions, velocities, discontinuity)

    def set_motion_ids(self, ids, n: int, positions, velocities, discontinuity=None) -> None:
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(n, 3)
        v = np.ascontiguousarray(velocities, dtype=np.float32).reshape(n, 3)
        d = None
        if discontinuity is not None:
            d = np.ascontiguousarray(discontinuity, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
        check(_lib.load().odb_spatial_set_motion_many(self._scene._h, n, ids, _fptr(p), _fptr(v), d))

    def set_listener_rotation(self, rotation_xyzs) -> None:
        """spatial.rs:345-349; mint::Quaternion as (x, y, z, s)."""
        q = np.ascontiguousarray(rotation_xyzs, dtype=np.float32).reshape(-1)
        if q.shape[0] != 4:
            raise ValueError("expected a quaternion (x, y, z, s)")
        check(_lib.load().odb_scene_set_listener_rotation(self._scene._h, _fptr(q)))


class Mixed:
    """mixer.rs:30-44"""

    def __init__(self, mixer_h, src: int):
        self._mixer, self._src = mixer_h, src

    def stop(self) -> None:
        check(_lib.load().odb_mixed_stop(self._mixer, self._src))

    def is_stopped(self) -> bool:
        out = C.c_int()
        check(_lib.load().odb_mixed_is_stopped(self._mixer, self._src, C.byref(out)))
        return bool(out.value)


class Mixer(_Aggregator):
    """mixer.rs:61-120"""

    def __init__(self, ctx: Context, channels: int):
        L = _lib.load()
        h = C.c_void_p()
        check(L.odb_mixer_create(ctx._h, int(channels), C.byref(h)))
        self._h, self._ctx, self.channels = h, ctx, int(channels)
        self._sample, self._sample_device, self._sample_i16 = L.odb_mixer_sample, L.odb_mixer_sample_device, L.odb_mixer_sample_i16
        self._destroy, self._set_epilogue = L.odb_mixer_destroy, L.odb_mixer_set_epilogue

    @staticmethod
    def new(channels: int = 2, ctx: Optional[Context] = None) -> Tuple["MixerControl", "Mixer"]:
        """Mixer::new (mixer.rs:70-81) -> (control, signal); channels stands in for the Frame type T."""
        m = Mixer(ctx or default_context(), channels)
        return MixerControl(m), m

    def __len__(self) -> int:
        out = C.c_uint64()
        check(_lib.load().odb_mixer_len(self._h, C.byref(out)))
        return out.value


class MixerControl:
    """mixer.rs:8-27"""

    def __init__(self, mixer: Mixer):
        self._mixer = mixer

    def play(self, signal: Signal) -> Mixed:
        if signal.channels != self._mixer.channels:
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "signal Frame type differs from the mixer's")
        chain, controls = _build_chain(signal)
        src = C.c_uint64()
        check(_lib.load().odb_mixer_play(self._mixer._h, C.byref(chain), C.byref(src)))
        for c in controls:
            c._bind(self._mixer._h, src.value)
        return Mixed(self._mixer._h, src.value)


class _Epilogue(Signal):
    """Tanh<T> / Reinhard<T> around an aggregator: fused into the reduce kernel's epilogue."""

    _code = _lib.EPILOGUE_NONE

    def __init__(self, inner: _Aggregator):
        if not isinstance(inner, _Aggregator):
            raise OddioError(_lib.ODB_E_UNSUPPORTED, "device path: Tanh/Reinhard wrap a SpatialScene or Mixer")
        self.inner, self.channels = inner, inner.channels
        check(inner._set_epilogue(inner._h, self._code))

    def sample(self, interval: float, n_or_out) -> np.ndarray:
        return self.inner.sample(interval, n_or_out)

    def sample_device(self, interval: float, dev_ptr: int, n_frames: int) -> None:
        self.inner.sample_device(interval, dev_ptr, n_frames)

    def sample_i16(self, interval: float, n_frames: int) -> np.ndarray:
        return self.inner.sample_i16(interval, n_frames)

    def is_finished(self) -> bool:
        return self.inner.is_finished()


class Tanh(_Epilogue):
    """tanh.rs:7-44"""

    _code = _lib.EPILOGUE_TANH


class Reinhard(_Epilogue):
    """reinhard.rs:13-50"""

    _code = _lib.EPILOGUE_REINHARD


# ---------------------------------------------------------------------------------------------
def run(signal, sample_rate: int, out) -> np.ndarray:
    """oddio::run (lib.rs:90-93): interval = 1.0 / sample_rate as f32; signal.sample(interval, out)."""
    interval = np.float32(1.0) / np.float32(sample_rate)
    return signal.sample(float(interval), out)


def frame_stereo(xs: np.ndarray) -> np.ndarray:
    """lib.rs:98-100: view a flat f32 buffer as stereo frames."""
    return xs.reshape(-1, 2)


def flatten_stereo(xs: np.ndarray) -> np.ndarray:
    """lib.rs:102-104"""
    return xs.reshape(-1)
