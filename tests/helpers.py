"""Builds the same scene / mixer twice — in the CPU oracle and on the device through the C ABI —
from one seeded description, so parity tests can compare them callback by callback."""
from __future__ import annotations

import numpy as np

F32 = np.float32


def synth_pcm(rng: np.random.Generator, n: int, rate: int, channels: int = 1) -> np.ndarray:
    """Sine + noise, like SURVEY.md §8d's synthetic PCM."""
    k = np.arange(n, dtype=np.float64)
    out = []
    for _ in range(channels):
        f = rng.uniform(100.0, 4000.0)
        ph = rng.uniform(0.0, 2 * np.pi)
        out.append(0.5 * np.sin(2 * np.pi * f * k / rate + ph) + 0.05 * rng.uniform(-1, 1, n))
    a = np.stack(out, axis=1).astype(F32)
    return a[:, 0].copy() if channels == 1 else a


def rand_in_shell(rng, r0, r1):
    v = rng.normal(size=3)
    v /= np.linalg.norm(v)
    return (v * rng.uniform(r0, r1)).astype(F32)


class ScenePair:
    """An oracle SpatialScene and a device SpatialScene fed identical calls."""

    def __init__(self, oracle, odb, ctx):
        self.o, self.odb = oracle, odb
        self.ref = oracle.SpatialScene()
        self.ctl, self.dev = odb.SpatialScene.new(ctx)
        self.ctx = ctx
        self.ref_handles, self.dev_handles = [], []
        self.ref_signals, self.dev_controls = [], []
        self._frames_cache = {}

    def frames(self, rate, pcm):
        key = id(pcm)
        if key not in self._frames_cache:
            self._frames_cache[key] = (self.o.Frames.from_slice(rate, pcm), self.odb.Frames.from_slice(rate, pcm, self.ctx), pcm)
        return self._frames_cache[key][:2]

    def play(self, rate, pcm, start, pos, vel, radius=0.1, fixed_gain_db=None):
        fo, fd = self.frames(rate, pcm)
        so = self.o.FramesSignal(fo, start)
        cd, sd = self.odb.FramesSignal.new(fd, start)
        inner_o, inner_d = so, sd
        if fixed_gain_db is not None:
            inner_o = self.o.FixedGain(so, fixed_gain_db)
            inner_d = self.odb.FixedGain(sd, fixed_gain_db)
        self.ref_handles.append(self.ref.play(inner_o, pos, vel, radius))
        self.dev_handles.append(self.ctl.play(inner_d, self.odb.SpatialOptions(pos, vel, radius)))
        self.ref_signals.append(so)
        self.dev_controls.append(cd)
        return len(self.ref_handles) - 1

    def play_buffered(self, rate, pcm, start, pos, vel, radius, max_distance, ring_rate, buffer_duration, speed=None,
                      gain=None, fixed_gain_db=None):
        fo, fd = self.frames(rate, pcm)
        so = self.o.FramesSignal(fo, start)
        cd, sd = self.odb.FramesSignal.new(fd, start)
        io, idv = so, sd
        ctl = {}
        if speed is not None:
            io = self.o.Speed(io); io.set_speed(speed)
            sc, idv = self.odb.Speed.new(idv); sc.set_speed(speed)
            ctl["speed"] = (io, sc)
        if fixed_gain_db is not None:
            io = self.o.FixedGain(io, fixed_gain_db)
            idv = self.odb.FixedGain(idv, fixed_gain_db)
        if gain is not None:
            io = self.o.Gain(io); io.set_amplitude_ratio(gain)
            gc, idv = self.odb.Gain.new(idv); idv.set_amplitude_ratio(gain)
            ctl["gain"] = (io, gc)
        self.ref_handles.append(self.ref.play_buffered(io, pos, vel, radius, max_distance, ring_rate, buffer_duration))
        self.dev_handles.append(self.ctl.play_buffered(idv, self.odb.SpatialOptions(pos, vel, radius), max_distance, ring_rate,
                                                       buffer_duration))
        self.ref_signals.append(so)
        self.dev_controls.append(cd)
        return len(self.ref_handles) - 1, ctl

    def set_motion(self, i, pos, vel, disc):
        self.ref_handles[i].set_motion(pos, vel, disc)
        self.dev_handles[i].set_motion(pos, vel, disc)

    def set_listener_rotation(self, q):
        self.ref.set_listener_rotation(q)
        self.ctl.set_listener_rotation(q)

    def step(self, sample_rate, n):
        """One oddio::run callback on both. Returns (ref f32, ref f64-accumulated, device)."""
        ref = self.o.run(self.ref, sample_rate, n)
        ref64 = self.ref.out64(n)
        out = np.zeros((n, 2), dtype=F32)
        self.odb.run(self.dev, sample_rate, out)
        return ref, ref64, out


def assert_mix_close(dev, ref, ref64, rel=1e-5):
    """SURVEY.md §7 H4 protocol: per sample within rel * max(|ref|, RMS of the buffer) of the
    reference-order f32 sum, and the device's error against the f64-accumulated truth is bounded
    the same way."""
    ref = np.asarray(ref, dtype=np.float64)
    dev = np.asarray(dev, dtype=np.float64)
    rms = float(np.sqrt(np.mean(ref64 ** 2))) if ref64.size else 0.0
    tol = rel * np.maximum(np.abs(ref), rms) + 1e-30
    bad = np.abs(dev - ref) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} samples off; worst {np.max(np.abs(dev - ref) / tol):.3g}x tolerance"
    tol64 = rel * np.maximum(np.abs(ref64), rms) + 1e-30
    bad64 = np.abs(dev - ref64) > tol64
    assert not bad64.any(), f"{bad64.sum()} samples off against the f64 truth"


class MixerPair:
    """An oracle Mixer and a device Mixer fed identical calls. `epilogue` in (None, 'tanh', 'reinhard')."""

    def __init__(self, oracle, odb, ctx, channels, epilogue=None):
        self.o, self.odb, self.ctx, self.channels = oracle, odb, ctx, channels
        self.ref_mixer = oracle.Mixer(channels)
        self.ctl, self.dev_mixer = odb.Mixer.new(channels, ctx)
        self.ref, self.dev = self.ref_mixer, self.dev_mixer
        if epilogue == "tanh":
            self.ref, self.dev = oracle.Tanh(self.ref_mixer), odb.Tanh(self.dev_mixer)
        elif epilogue == "reinhard":
            self.ref, self.dev = oracle.Reinhard(self.ref_mixer), odb.Reinhard(self.dev_mixer)
        self.items = []  # dicts: ref/dev handles and controls per played signal
        self._frames_cache = {}

    def frames(self, rate, pcm):
        key = id(pcm)
        if key not in self._frames_cache:
            self._frames_cache[key] = (self.o.Frames.from_slice(rate, pcm), self.odb.Frames.from_slice(rate, pcm, self.ctx), pcm)
        return self._frames_cache[key][:2]

    def play(self, rate, pcm, start=0.0, speed=None, fixed_gain_db=None, gain=None):
        fo, fd = self.frames(rate, pcm)
        so = self.o.FramesSignal(fo, start)
        cd, sd = self.odb.FramesSignal.new(fd, start)
        it = {"ref_frames_signal": so, "dev_frames_control": cd}
        io, idv = so, sd
        if speed is not None:
            io = self.o.Speed(io); io.set_speed(speed)
            sc, idv = self.odb.Speed.new(idv); sc.set_speed(speed)
            it["ref_speed"], it["dev_speed"] = io, sc
        if fixed_gain_db is not None:
            io = self.o.FixedGain(io, fixed_gain_db)
            idv = self.odb.FixedGain(idv, fixed_gain_db)
        if gain is not None:
            io = self.o.Gain(io); io.set_amplitude_ratio(gain)
            gc, idv = self.odb.Gain.new(idv); idv.set_amplitude_ratio(gain)
            it["ref_gain"], it["dev_gain"] = io, gc
        it["ref_mixed"] = self.ref_mixer.play(io)
        it["dev_mixed"] = self.ctl.play(idv)
        self.items.append(it)
        return len(self.items) - 1

    def set_speed(self, i, v):
        self.items[i]["ref_speed"].set_speed(v)
        self.items[i]["dev_speed"].set_speed(v)

    def set_gain_ratio(self, i, v):
        self.items[i]["ref_gain"].control_set_amplitude_ratio(v)
        self.items[i]["dev_gain"].set_amplitude_ratio(v)

    def stop(self, i):
        self.items[i]["ref_mixed"].stop()
        self.items[i]["dev_mixed"].stop()

    def step(self, sample_rate, n):
        ref = self.o.run(self.ref, sample_rate, n)
        ref64 = self.ref_mixer.out64(n)
        shape = (n, self.channels) if self.channels > 1 else (n,)
        out = np.zeros(shape, dtype=F32)
        self.odb.run(self.dev, sample_rate, out)
        return ref, ref64, out
