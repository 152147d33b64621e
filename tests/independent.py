"""(Test infrastructure.) A second, independent statement of the hot path - SpatialScene::sample (seek and buffered
sets) and Mixer::sample over its chains - to hold the C++ oracle and the committed golden vectors against (CPU only).

The reference holds no numeric test of `SpatialScene::sample` (spatial.rs:606-665 only checks when a finished signal is
dropped) and none of a whole `Mixer` chain, so the oracle's numbers for the headline path rest on how faithfully it
restates the source. This file restates the same lines a second time, in another language and
written from the reference alone - scalar numpy float32 / float64 operations in the reference's order;
tests/test_oracle_independent.py and tests/test_golden.py require the C++ oracle and the golden vectors to agree with
it BIT FOR BIT on output, f64 cursors and removal behaviour. Two independent restatements agreeing is not the reference itself, but
it rules out the slips a single restatement can hide (operation order, f32 vs f64 intermediates, truncation, chunking).

Reference lines: lib.rs:90-93 (run), spatial.rs:191-265 (walk_set), 345-349 (set_listener_rotation), 376-471 (sample,
seek set), 489-503 (smoothed_position), 522-543 (EarState::new), 563-598 (Ear), math/mod.rs:32-99, frames.rs:105-123
(get_pair), 176-213 (FramesSignal), frame.rs:39-41 (lerp); buffered set: spatial.rs:30-57, 313-340, 395-433,
ring.rs:9-79; mixer: mixer.rs:77, 92-119, gain.rs:27-43, 58-127, 163, smooth.rs:26-91, speed.rs:24-40; the rest of the
closed set: cycle.rs:6-61, sine.rs:6-46, signal.rs:62-86 (MonoToStereo), reinhard.rs:28-35, tanh.rs:24-28. Not restated:
libm: `powf` (FixedGain's ratio), `tanhf` (the Tanh epilogue) and `sinf` (Sine) are glibc's, called through ctypes - the same
library the oracle links and a Rust std build would call.

(Writing it caught a slip - in THIS file: the first draft forgot that `set_listener_rotation` stores the inverse
rotation; the C++ oracle had it right.)"""
import ctypes

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.powf.restype = _libm.tanhf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.tanhf.argtypes = [ctypes.c_float]

f32 = np.float32
f64 = np.float64

SPEED_OF_SOUND = f32(343.0)
HEAD_RADIUS = f32(0.1075)
POSITION_SMOOTHING_PERIOD = f32(0.5)
EPSILON = np.finfo(np.float32).eps


def v3(x):
    return [f32(x[0]), f32(x[1]), f32(x[2])]


def norm(x):  # math/mod.rs:32-34
    s = f32(0.0)
    for c in x:
        s = f32(s + f32(c * c))
    return f32(np.sqrt(s))


def dot(x, y):  # :36-42
    s = f32(0.0)
    for a, b in zip(x, y):
        s = f32(s + f32(a * b))
    return s


def scale(v, k):  # :44-46
    return [f32(v[0] * k), f32(v[1] * k), f32(v[2] * k)]


def sub(a, b):  # :48-50
    return [f32(a[0] - b[0]), f32(a[1] - b[1]), f32(a[2] - b[2])]


def add(a, b):  # :52-54
    return [f32(a[0] + b[0]), f32(a[1] + b[1]), f32(a[2] + b[2])]


def mix(a, b, r):  # :56-59
    ir = f32(f32(1.0) - r)
    return [f32(f32(ir * a[i]) + f32(r * b[i])) for i in range(3)]


def quat_mul(q, r):  # :68-79; q = (s, [x, y, z])
    qs, (qx, qy, qz) = q
    rs, (rx, ry, rz) = r

    def chain(a, b, c, d):  # a + b + c + d evaluated left to right, each term already rounded
        return f32(f32(f32(a + b) + c) + d)

    s = chain(f32(qs * rs), -f32(qx * rx), -f32(qy * ry), -f32(qz * rz))
    x = chain(f32(qs * rx), f32(qx * rs), f32(qy * rz), -f32(qz * ry))
    y = chain(f32(qs * ry), -f32(qx * rz), f32(qy * rs), f32(qz * rx))
    z = chain(f32(qs * rz), f32(qx * ry), -f32(qy * rx), f32(qz * rs))
    return (s, [x, y, z])


def rotate(rot, p):  # :81-95
    inv = (rot[0], [f32(-rot[1][0]), f32(-rot[1][1]), f32(-rot[1][2])])
    return quat_mul(rot, quat_mul((f32(0.0), list(p)), inv))[1]


SQRT17 = f32(np.sqrt(f32(17.0)))


def ear_pos(ear):  # spatial.rs:565-576
    return [f32(-HEAD_RADIUS) if ear == 0 else HEAD_RADIUS, f32(0.0), f32(0.0)]


def ear_dir(ear):  # :579-597: [+-4, 0, -1] normalised
    sign = f32(-1.0) if ear == 0 else f32(1.0)
    return [f32(f32(sign * f32(4.0)) / SQRT17), f32(0.0), f32(f32(-1.0) / SQRT17)]


def ear_state(p, ear, radius):  # :522-543 -> (offset, gain)
    distance = norm(sub(p, ear_pos(ear)))
    offset = f32(distance * f32(f32(-1.0) / SPEED_OF_SOUND))
    distance_gain = f32(radius / max(distance, radius))
    if distance < f32(1e-3):
        stereo = f32(f32(0.5) + f32(0.5))
    else:
        stereo = f32(f32(0.5) + dot(ear_dir(ear), scale(p, f32(f32(0.5) / distance))))
    return offset, f32(stereo * distance_gain)


def trunc_isize(x):
    return int(np.trunc(x))


class PyFramesSignal:
    """FramesSignal<T> for T = f32 (samples of shape (n,)) or T = [f32; ch] (samples of shape (n, ch): elementwise f32
    array arithmetic rounds every channel like the scalar operation, so the lerp lines below serve both)."""

    def __init__(self, samples, rate, start_seconds):
        self.samples, self.rate, self.t = np.asarray(samples, dtype=f32), f64(rate), f64(start_seconds)

    def get_pair(self, s):  # frames.rs:105-123
        n = self.samples.shape[0]
        z = f32(0.0) if self.samples.ndim == 1 else np.zeros(self.samples.shape[1], dtype=f32)
        if s >= 0:
            if s < n - 1:
                return self.samples[s], self.samples[s + 1]
            if s < n:
                return self.samples[s], z
            return z, z
        if s < -1:
            return z, z
        return z, self.samples[0]

    def sample(self, interval, n):  # :176-201
        out = np.empty((n,) + self.samples.shape[1:], dtype=f32)
        s0 = f64(self.t * self.rate)
        ds = f32(interval * f32(self.rate))
        base = trunc_isize(s0)
        if abs(f32(ds - f32(1.0))) <= EPSILON:
            fract = f32(s0 - f64(base))
            for i in range(n):
                a, b = self.get_pair(base + i)
                out[i] = f32(a + f32(fract * f32(b - a)))  # frame.rs:39-41
        else:
            offset = f32(s0 - f64(base))
            for i in range(n):
                tr = trunc_isize(offset)
                a, b = self.get_pair(base + tr)
                fract = f32(offset - f32(tr))
                out[i] = f32(a + f32(fract * f32(b - a)))
                offset = f32(offset + ds)
        self.t = f64(self.t + f64(f64(interval) * f64(n)))
        return out

    def is_finished(self):  # :204-206
        return self.t >= f64(self.samples.shape[0] - 1) / self.rate

    def seek(self, seconds):  # :210-213
        self.t = f64(self.t + f64(f32(seconds)))


class PySource:
    def __init__(self, inner, position, velocity, radius):  # spatial.rs:60-117
        self.inner, self.radius = inner, f32(radius)
        self.received = (v3(position), v3(velocity), False)
        self.pending = None
        self.prev_position, self.dt = v3(position), f32(0.0)
        self.finished_for, self.stopped = None, False

    def set_motion(self, position, velocity, discontinuity):  # :137-149
        self.pending = (v3(position), v3(velocity), bool(discontinuity))

    def smoothed_position(self, dt, motion):  # :489-503
        dt = f32(self.dt + dt)
        change = scale(motion[1], dt)
        naive = add(self.prev_position, change)
        intended = add(motion[0], change)
        return mix(naive, intended, min(f32(dt / POSITION_SMOOTHING_PERIOD), f32(1.0)))


class PyScene:
    def __init__(self):
        self.sources, self.buffered = [], []  # the seek set and the buffered set
        self.rot_received = (f32(1.0), [f32(0.0)] * 3)
        self.rot_pending = None

    def play(self, inner, position, velocity, radius):
        self.sources.append(PySource(inner, position, velocity, radius))
        return self.sources[-1]

    def set_listener_rotation(self, xyzs):  # spatial.rs:345-349: the scene turns the other way
        self.rot_pending = (f32(xyzs[3]), [f32(-xyzs[0]), f32(-xyzs[1]), f32(-xyzs[2])])

    def run(self, sample_rate, n):  # lib.rs:90-93, spatial.rs:376-471
        interval = f32(f32(1.0) / f32(sample_rate))
        prev_rot = self.rot_received
        if self.rot_pending is not None:
            self.rot_received, self.rot_pending = self.rot_pending, None
        rot = self.rot_received
        out = np.zeros((n, 2), dtype=f32)
        elapsed = f32(interval * f32(n))
        self.walk(self.buffered, self.mix_buffered, prev_rot, rot, elapsed, n, out)  # :395-433
        self.walk(self.sources, self.mix, prev_rot, rot, elapsed, n, out)            # :435-470
        return out

    def walk(self, sources, mix_signal, prev_rot, rot, elapsed, n, out):  # walk_set, spatial.rs:191-265
        for i in reversed(range(len(sources))):
            src = sources[i]
            orig_next = src.received
            if src.pending is not None:  # motion.refresh(), :216-224
                src.received, src.pending = src.pending, None
                src.prev_position = src.received[0] if src.received[2] else src.smoothed_position(f32(0.0), orig_next)
                src.dt = f32(0.0)
            prev_position = rotate(prev_rot, src.smoothed_position(f32(0.0), src.received))
            next_position = rotate(rot, src.smoothed_position(elapsed, src.received))
            src.dt = f32(src.dt + elapsed)
            distance = norm(prev_position)  # :241-258
            if src.finished_for is not None:
                if src.finished_for > f32(distance / SPEED_OF_SOUND):
                    src.stopped = True
                else:
                    src.finished_for = f32(src.finished_for + elapsed)
            elif src.inner.is_finished():
                src.finished_for = elapsed
            if src.stopped:
                del sources[i]  # set.remove = swap_remove (set.rs:183-188); same thing at the walked index ...
                if i < len(sources):  # ... unless something sits behind it: the last element moves into the hole
                    sources.insert(i, sources.pop())
                continue
            mix_signal(src, prev_position, next_position, elapsed, n, out)

    def mix(self, src, prev_position, next_position, elapsed, n, out):  # the seek set's closure, spatial.rs:445-469
        for ear in (0, 1):
            p_off, p_gain = ear_state(prev_position, ear, src.radius)
            n_off, n_gain = ear_state(next_position, ear, src.radius)
            src.inner.seek(p_off)
            effective = f32(f32(elapsed + n_off) - p_off)
            dt = f32(effective / f32(n))
            d_gain = f32(f32(n_gain - p_gain) / f32(n))
            k = 0
            for c0 in range(0, n, 256):
                m = min(256, n - c0)
                buf = src.inner.sample(dt, m)
                for s in buf:
                    gain = f32(p_gain + f32(f32(k) * d_gain))
                    out[k, ear] = f32(out[k, ear] + f32(s * gain))
                    k += 1
            src.inner.seek(f32(f32(-effective) - p_off))
        src.inner.seek(elapsed)


def make_pcm(rng, n, rate):
    k = np.arange(n, dtype=np.float64)
    return (0.5 * np.sin(2 * np.pi * rng.uniform(100.0, 4000.0) * k / rate + rng.uniform(0, 6.28)) + 0.05 * rng.uniform(-1, 1, n)).astype(f32)




# ---- the mixer path: Mixer<T> over Gain(FixedGain(Speed(FramesSignal))) ------------------------------------------------
SMOOTHING_PERIOD = f32(0.1)  # gain.rs:163


class PySpeed:  # speed.rs:24-40
    def __init__(self, inner):
        self.inner, self.speed = inner, f32(1.0)

    def sample(self, interval, n):
        return self.inner.sample(f32(interval * self.speed), n)

    def is_finished(self):
        return self.inner.is_finished()


def db_to_ratio(db):  # gain.rs:20 `10.0f32.powf(db / 20.0)`
    return f32(_libm.powf(10.0, float(f32(f32(db) / f32(20.0)))))


def tanh32(x):  # tanh.rs:24-28, per channel
    flat = np.asarray(x, dtype=f32).ravel()
    return np.array([_libm.tanhf(float(v)) for v in flat], dtype=f32).reshape(np.shape(x))


class PyFixedGain:  # gain.rs:13-54
    def __init__(self, inner, ratio):
        self.inner, self.gain = inner, f32(ratio)

    def seek(self, seconds):
        self.inner.seek(seconds)

    def sample(self, interval, n):
        return (self.inner.sample(interval, n) * self.gain).astype(f32)

    def is_finished(self):
        return self.inner.is_finished()


class PyGain:  # gain.rs:58-127 over smooth.rs:26-91
    def __init__(self, inner):
        self.inner, self.shared = inner, f32(1.0)
        self.prev, self.next, self.progress = f32(1.0), f32(1.0), f32(1.0)

    def set_amplitude_ratio(self, factor):  # Gain::set_amplitude_ratio: no smoothing
        self.shared = f32(factor)
        self.prev, self.next, self.progress = f32(factor), f32(factor), f32(1.0)

    def control_set_amplitude_ratio(self, factor):  # GainControl::set_amplitude_ratio
        self.shared = f32(factor)

    def get(self):  # Smoothed::get -> f32::interpolate
        return f32(self.prev + f32(self.progress * f32(self.next - self.prev)))

    def sample(self, interval, n):
        out = self.inner.sample(interval, n)
        if self.next != self.shared:  # Smoothed::set
            self.prev, self.next, self.progress = self.get(), self.shared, f32(0.0)
        if self.progress == f32(1.0):
            g = self.get()
            if g != f32(1.0):
                out = (out * g).astype(f32)
            return out
        for i in range(n):
            out[i] = f32(out[i] * self.get()) if out.ndim == 1 else (out[i] * self.get()).astype(f32)
            self.progress = min(f32(self.progress + f32(interval / SMOOTHING_PERIOD)), f32(1.0))
        return out

    def is_finished(self):
        return self.inner.is_finished()


class PyMixer:  # mixer.rs:92-119
    def __init__(self, channels):
        self.channels, self.signals = channels, []

    def play(self, signal):
        entry = {"inner": signal, "stop": False}
        self.signals.append(entry)
        return entry

    def run(self, sample_rate, n):
        interval = f32(f32(1.0) / f32(sample_rate))
        out = np.zeros((n, self.channels) if self.channels > 1 else (n,), dtype=f32)
        for i in reversed(range(len(self.signals))):
            sig = self.signals[i]
            if sig["stop"] or sig["inner"].is_finished():
                sig["stop"] = True
                del self.signals[i]  # swap_remove
                if i < len(self.signals):
                    self.signals.insert(i, self.signals.pop())
                continue
            for c0 in range(0, n, 1024):  # the staging buffer holds 1024 frames (mixer.rs:77)
                m = min(1024, n - c0)
                out[c0:c0 + m] = (out[c0:c0 + m] + sig["inner"].sample(interval, m)).astype(f32)
        return out




# ---- the buffered path: SpatialSceneControl::play_buffered over Ring ------------------------------------------------
def fmod32(a, b):  # Rust's `%` on f32 is fmod: exact
    return f32(np.fmod(f32(a), f32(b)))


def rem_euclid32(a, b):  # f32::rem_euclid
    r = fmod32(a, b)
    return f32(r + abs(b)) if r < f32(0.0) else r


class PyRing:  # ring.rs:4-79
    def __init__(self, capacity):
        self.buffer, self.write = np.zeros(capacity, dtype=f32), f32(0.0)

    def write_from(self, signal, rate, dt):  # Ring::write
        n = f32(self.buffer.size)
        end = fmod32(f32(self.write + f32(dt * f32(rate))), n)
        start_idx, end_idx = int(np.ceil(self.write)), int(np.ceil(end))
        interval = f32(f32(1.0) / f32(rate))
        if end_idx > start_idx:
            self.buffer[start_idx:end_idx] = signal.sample(interval, end_idx - start_idx)
        else:
            self.buffer[start_idx:] = signal.sample(interval, self.buffer.size - start_idx)
            self.buffer[:end_idx] = signal.sample(interval, end_idx)
        self.write = end

    def delay(self, rate, dt):
        self.write = fmod32(f32(self.write + f32(f32(rate) * dt)), f32(self.buffer.size))

    def sample(self, rate, t, interval, n):
        size = self.buffer.size
        out = np.empty(n, dtype=f32)
        offset = rem_euclid32(f32(self.write + f32(t * f32(rate))), f32(size))
        ds = f32(interval * f32(rate))
        for i in range(n):
            trunc = int(np.trunc(offset))
            fract = f32(offset - f32(trunc))
            x = trunc
            if x < size - 1:
                a, b = self.buffer[x], self.buffer[x + 1]
            elif x < size:
                a, b = self.buffer[x], self.buffer[0]
            else:
                x = x % size
                offset = f32(f32(x) + fract)
                a, b = (self.buffer[x], self.buffer[x + 1]) if x < size - 1 else (self.buffer[x], self.buffer[0])
            out[i] = f32(a + f32(fract * f32(b - a)))
            offset = f32(offset + ds)
        return out


def play_buffered(self, inner, position, velocity, radius, max_distance, rate, buffer_duration):  # spatial.rs:313-340, :30-57
    src = PySource(inner, position, velocity, radius)
    src.rate = int(rate)
    src.max_delay = f32(f32(f32(max_distance) / SPEED_OF_SOUND) + f32(buffer_duration))
    src.queue = PyRing(int(np.ceil(f32(src.max_delay * f32(rate)))) + 1)
    src.queue.delay(rate, min(f32(norm(v3(position)) / SPEED_OF_SOUND), src.max_delay))
    self.buffered.append(src)
    return src


def mix_buffered(self, src, prev_position, next_position, elapsed, n, out):  # the buffered set's closure, spatial.rs:404-432
    src.queue.write_from(src.inner, src.rate, elapsed)
    for ear in (0, 1):
        p_off, p_gain = ear_state(prev_position, ear, src.radius)
        n_off, n_gain = ear_state(next_position, ear, src.radius)
        prev_offset = max(f32(p_off - elapsed), f32(-src.max_delay))
        next_offset = max(n_off, f32(-src.max_delay))
        dt = f32(f32(next_offset - prev_offset) / f32(n))
        d_gain = f32(f32(n_gain - p_gain) / f32(n))
        k = 0
        for c0 in range(0, n, 256):
            m = min(256, n - c0)
            t = f32(prev_offset + f32(f32(k) * dt))
            buf = src.queue.sample(src.rate, t, dt, m)
            for s in buf:
                gain = f32(p_gain + f32(f32(k) * d_gain))
                out[k, ear] = f32(out[k, ear] + f32(s * gain))
                k += 1


PyScene.play_buffered = play_buffered
PyScene.mix_buffered = mix_buffered




# ---- the backend interface of tests/golden/scenarios.py -------------------------------------------------------------------
class IndependentBackend:
    """Drives the restatement through the calls of a golden scenario (like OracleBackend / DeviceBackend there)."""

    def frames(self, rate, pcm):
        return (rate, np.asarray(pcm, dtype=f32))

    def scene(self):
        sc, sigs = PyScene(), []

        class S:
            def play(_, fr, start, pos, vel, radius=0.1, fixed_gain_db=None):
                s = PyFramesSignal(fr[1], fr[0], start)
                sigs.append(s)
                inner = s if fixed_gain_db is None else PyFixedGain(s, db_to_ratio(fixed_gain_db))
                return sc.play(inner, pos, vel, radius)

            def play_buffered(_, fr, start, pos, vel, radius, max_distance, rate, buffer_duration, gain=None):
                s = PyFramesSignal(fr[1], fr[0], start)
                sigs.append(s)
                inner = s
                if gain is not None:
                    inner = PyGain(s)
                    inner.set_amplitude_ratio(gain)
                return sc.play_buffered(inner, pos, vel, radius, max_distance, rate, buffer_duration)

            def set_listener_rotation(_, q):
                sc.set_listener_rotation(q)

            def run(_, rate, n):
                return sc.run(rate, n)

            def cursors(_):
                return np.array([float(s.t) for s in sigs], dtype=np.float64)

        return S()

    def mixer(self, channels, tanh=False):
        mx, sigs = PyMixer(channels), []

        class M:
            def play(_, fr, start, speed=None, gain=None, fixed_gain_db=None):
                s = PyFramesSignal(fr[1], fr[0], start)
                sigs.append(s)
                inner = s
                if speed is not None:
                    inner = PySpeed(inner)
                    inner.speed = f32(speed)
                if fixed_gain_db is not None:
                    inner = PyFixedGain(inner, db_to_ratio(fixed_gain_db))
                g = None
                if gain is not None:
                    inner = g = PyGain(inner)
                    g.set_amplitude_ratio(gain)
                mx.play(inner)
                return (lambda v: g.control_set_amplitude_ratio(v)) if g is not None else None

            def run(_, rate, n):
                out = mx.run(rate, n)
                return tanh32(out) if tanh else out

            def cursors(_):
                return np.array([float(s.t) for s in sigs], dtype=np.float64)

        return M()


# ---- the rest of the closed set: Cycle, Sine, MonoToStereo, Reinhard -------------------------------------------------
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
TAU = f32(6.28318530717958647692528676655900577)  # core::f32::consts::TAU


class PyCycle:  # cycle.rs:6-61
    def __init__(self, samples, rate):
        self.samples, self.rate, self.cursor = np.asarray(samples, dtype=f32), int(rate), f64(0.0)

    def sample(self, interval, n):
        size = self.samples.shape[0]
        out = np.empty((n,) + self.samples.shape[1:], dtype=f32)
        ds = f32(interval * f32(self.rate))
        base = int(np.trunc(self.cursor))
        offset = f32(self.cursor - f64(base))
        for i in range(n):
            trunc = int(np.trunc(offset))
            fract = f32(offset - f32(trunc))
            x = base + trunc
            if x < size - 1:
                a, b = self.samples[x], self.samples[x + 1]
            elif x < size:
                a, b = self.samples[x], self.samples[0]
            else:
                base = 0
                offset = f32(f32(x % size) + fract)
                x = int(np.trunc(offset))
                a, b = (self.samples[x], self.samples[x + 1]) if x < size - 1 else (self.samples[x], self.samples[0])
            out[i] = f32(a + f32(fract * f32(b - a)))
            offset = f32(offset + ds)
        self.cursor = f64(f64(base) + f64(offset))
        return out

    def is_finished(self):  # Signal's default
        return False

    def seek(self, seconds):
        size = f64(self.samples.shape[0])
        r = np.fmod(f64(self.cursor + f64(f64(f32(seconds)) * f64(self.rate))), size)
        self.cursor = f64(r + size) if r < 0 else f64(r)


class PySine:  # sine.rs:6-46
    def __init__(self, phase, frequency_hz):
        self.phase, self.frequency = f32(phase), f32(f32(frequency_hz) * TAU)

    def seek(self, t):
        self.phase = f32(np.fmod(f32(self.phase + f32(f32(t) * self.frequency)), TAU))

    def sample(self, interval, n):
        out = np.empty(n, dtype=f32)
        for i in range(n):
            t = f32(interval * f32(i))
            out[i] = _libm.sinf(float(f32(f32(t * self.frequency) + self.phase)))
        self.seek(f32(interval * f32(n)))
        return out

    def is_finished(self):
        return False


class PyMonoToStereo:  # signal.rs:62-86
    def __init__(self, inner):
        self.inner = inner

    def sample(self, interval, n):
        mono = self.inner.sample(interval, n)
        return np.stack([mono, mono], axis=1)

    def is_finished(self):
        return self.inner.is_finished()

    def seek(self, seconds):
        self.inner.seek(seconds)


def reinhard32(x):  # reinhard.rs:28-35: channel /= 1 + |channel|
    x = np.asarray(x, dtype=f32)
    return (x / (f32(1.0) + np.abs(x))).astype(f32)
