"""Seeded scenarios behind the committed golden vectors (tests/golden/*.npz). Each scenario drives a
backend - the CPU oracle or the device path - through the same calls and returns the callback outputs plus
the per-source f64 cursors. make_golden.py records the oracle's results; the tests replay them."""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _pcm(rng, n, rate, ch=1):
    k = np.arange(n, dtype=np.float64)
    cols = []
    for _ in range(ch):
        f, ph = rng.uniform(100.0, 4000.0), rng.uniform(0, 2 * np.pi)
        cols.append(0.5 * np.sin(2 * np.pi * f * k / rate + ph) + 0.05 * rng.uniform(-1, 1, n))
    a = np.stack(cols, axis=1).astype(F32)
    return a[:, 0].copy() if ch == 1 else a


def _shell(rng, r0, r1):
    v = rng.normal(size=3)
    return (v / np.linalg.norm(v) * rng.uniform(r0, r1)).astype(F32)


class OracleBackend:
    """The CPU oracle (oracle/pyoracle.py)."""

    def __init__(self):
        from oracle import pyoracle

        self.o = pyoracle

    def frames(self, rate, pcm):
        return self.o.Frames.from_slice(rate, pcm)

    def scene(self):
        sc = self.o.SpatialScene()
        sigs = []

        class S:
            def play(_, fr, start, pos, vel, radius=0.1, fixed_gain_db=None):
                s = self.o.FramesSignal(fr, start)
                sigs.append(s)
                inner = s if fixed_gain_db is None else self.o.FixedGain(s, fixed_gain_db)
                return sc.play(inner, pos, vel, radius)

            def play_buffered(_, fr, start, pos, vel, radius, max_distance, rate, buffer_duration, gain=None):
                s = self.o.FramesSignal(fr, start)
                sigs.append(s)
                inner = s
                if gain is not None:
                    inner = self.o.Gain(s)
                    inner.set_amplitude_ratio(gain)
                return sc.play_buffered(inner, pos, vel, radius, max_distance, rate, buffer_duration)

            def set_listener_rotation(_, q):
                sc.set_listener_rotation(q)

            def run(_, rate, n):
                return self.o.run(sc, rate, n)

            def cursors(_):
                return np.array([s.t for s in sigs], dtype=np.float64)

        return S()

    def mixer(self, channels, tanh=False):
        mx = self.o.Mixer(channels)
        top = self.o.Tanh(mx) if tanh else mx
        sigs = []

        class M:
            def play(_, fr, start, speed=None, gain=None, fixed_gain_db=None):
                s = self.o.FramesSignal(fr, start)
                sigs.append(s)
                inner = s
                if speed is not None:
                    inner = self.o.Speed(inner)
                    inner.set_speed(speed)
                if fixed_gain_db is not None:
                    inner = self.o.FixedGain(inner, fixed_gain_db)
                g = None
                if gain is not None:
                    inner = g = self.o.Gain(inner)
                    g.set_amplitude_ratio(gain)
                mx.play(inner)
                return (lambda v: g.control_set_amplitude_ratio(v)) if g is not None else None

            def run(_, rate, n):
                return self.o.run(top, rate, n)

            def cursors(_):
                return np.array([s.t for s in sigs], dtype=np.float64)

        return M()


class DeviceBackend:
    """The device path through the C ABI (oddio_b200)."""

    def __init__(self, ctx=None):
        import oddio_b200 as odb

        self.odb = odb
        self.ctx = ctx or odb.init(0)

    def frames(self, rate, pcm):
        return self.odb.Frames.from_slice(rate, pcm, self.ctx)

    def scene(self):
        odb = self.odb
        ctl, sc = odb.SpatialScene.new(self.ctx)
        ctrls = []

        class S:
            def play(_, fr, start, pos, vel, radius=0.1, fixed_gain_db=None):
                c, s = odb.FramesSignal.new(fr, start)
                ctrls.append(c)
                inner = s if fixed_gain_db is None else odb.FixedGain(s, fixed_gain_db)
                return ctl.play(inner, odb.SpatialOptions(pos, vel, radius))

            def play_buffered(_, fr, start, pos, vel, radius, max_distance, rate, buffer_duration, gain=None):
                c, s = odb.FramesSignal.new(fr, start)
                ctrls.append(c)
                inner = s
                if gain is not None:
                    inner = odb.Gain(s)
                    inner.set_amplitude_ratio(gain)
                return ctl.play_buffered(inner, odb.SpatialOptions(pos, vel, radius), max_distance, rate, buffer_duration)

            def set_listener_rotation(_, q):
                ctl.set_listener_rotation(q)

            def run(_, rate, n):
                out = np.zeros((n, 2), dtype=F32)
                odb.run(sc, rate, out)
                return out

            def cursors(_):
                return np.array([c.cursor()[0] for c in ctrls], dtype=np.float64)

        return S()

    def mixer(self, channels, tanh=False):
        odb = self.odb
        ctl, mx = odb.Mixer.new(channels, self.ctx)
        top = odb.Tanh(mx) if tanh else mx
        ctrls = []

        class M:
            def play(_, fr, start, speed=None, gain=None, fixed_gain_db=None):
                c, s = odb.FramesSignal.new(fr, start)
                ctrls.append(c)
                inner = s
                if speed is not None:
                    sc_, inner = odb.Speed.new(inner)
                    sc_.set_speed(speed)
                if fixed_gain_db is not None:
                    inner = odb.FixedGain(inner, fixed_gain_db)
                gc = None
                if gain is not None:
                    gc, inner = odb.Gain.new(inner)
                    inner.set_amplitude_ratio(gain)
                ctl.play(inner)
                return (lambda v: gc.set_amplitude_ratio(v)) if gc is not None else None

            def run(_, rate, n):
                out = np.zeros((n, channels) if channels > 1 else (n,), dtype=F32)
                odb.run(top, rate, out)
                return out

            def cursors(_):
                return np.array([c.cursor()[0] for c in ctrls], dtype=np.float64)

        return M()


# ---- scenarios: name -> function(backend) -> dict of arrays ------------------------------------------------
def scene_seek(b):
    """24 moving sources + 2 static ones (ds ~= 1 path), motion update, listener rotation; 3 callbacks."""
    rng = np.random.default_rng(1001)
    rate = 48000
    pcms = [b.frames(rate, _pcm(rng, 56000, rate)) for _ in range(4)]
    sc = b.scene()
    hs = [sc.play(pcms[i % 4], 1.0, _shell(rng, 2, 100), rng.uniform(-30, 30, 3).astype(F32)) for i in range(24)]
    hs.append(sc.play(pcms[0], 0.25, [0.0, 0.0, -5.0], [0.0, 0.0, 0.0]))
    hs.append(sc.play(pcms[1], 0.5, [3.0, 0.0, 2.0], [0.0, 0.0, 0.0], fixed_gain_db=-6.0))
    out = {}
    for k, n in enumerate((256, 1024, 700)):
        if k == 1:
            hs[3].set_motion([5.0, 1.0, 0.0], [0.0, 2.0, 0.0], False)
            hs[7].set_motion([-8.0, 0.0, 4.0], [1.0, 0.0, 0.0], True)
            sc.set_listener_rotation(np.array([0.0, 0.3826834, 0.0, 0.9238795], dtype=F32))
        out[f"out{k}"] = sc.run(rate, n)
        out[f"t{k}"] = sc.cursors()
    return out


def scene_buffered(b):
    """6 buffered sources (one under Gain) + 4 seek sources; 4 callbacks (the ring wraps)."""
    rng = np.random.default_rng(1002)
    rate = 48000
    pcms = [b.frames(rate, _pcm(rng, 60000, rate)) for _ in range(3)]
    sc = b.scene()
    for i in range(6):
        sc.play_buffered(pcms[i % 3], 0.0, _shell(rng, 1, 40), rng.uniform(-10, 10, 3).astype(F32), 0.1, 60.0, rate, 0.05,
                         gain=0.5 if i == 2 else None)
    for i in range(4):
        sc.play(pcms[i % 3], 1.0, _shell(rng, 2, 40), rng.uniform(-10, 10, 3).astype(F32))
    out = {}
    for k, n in enumerate((512, 1024, 1024, 300)):
        out[f"out{k}"] = sc.run(rate, n)
        out[f"t{k}"] = sc.cursors()
    return out


def mixer_stereo_tanh(b):
    """C4 in small: 40 static stereo sources under Gain, Tanh over the mixer, 96 kHz; a gain transition."""
    rng = np.random.default_rng(1003)
    rate = 96000
    pcms = [b.frames(rate, _pcm(rng, 9000, rate, 2)) for _ in range(4)]
    mx = b.mixer(2, tanh=True)
    setters = [mx.play(pcms[i % 4], 0.0, gain=float(rng.uniform(0.05, 0.3))) for i in range(40)]
    out = {}
    for k in range(3):
        if k == 1:
            setters[0](0.9)
            setters[5](0.01)
        out[f"out{k}"] = mx.run(rate, 1024)
        out[f"t{k}"] = mx.cursors()
    return out


def mixer_speed_sweep(b):
    """C5 in small: 16 mono Speed<FramesSignal> sources, ratio 0.5..2.0, one 4096-frame callback + a short one."""
    rng = np.random.default_rng(1004)
    rate = 48000
    pcms = [b.frames(rate, _pcm(rng, 24000, rate)) for _ in range(2)]
    mx = b.mixer(1)
    for i in range(16):
        mx.play(pcms[i % 2], 0.0, speed=float(rng.uniform(0.5, 2.0)), fixed_gain_db=-3.0 if i % 4 == 0 else None)
    out = {}
    for k, n in enumerate((4096, 333)):
        out[f"out{k}"] = mx.run(rate, n)
        out[f"t{k}"] = mx.cursors()
    return out


SCENARIOS = {"scene_seek": scene_seek, "scene_buffered": scene_buffered, "mixer_stereo_tanh": mixer_stereo_tanh,
             "mixer_speed_sweep": mixer_speed_sweep}
