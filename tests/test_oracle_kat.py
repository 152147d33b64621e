"""Pins the CPU oracle against every known-answer vector the reference's unit tests hold for
the hot path. Each test cites the reference test it replays; values are the
reference's own `assert_eq!` literals (bit-exact, no tolerance) unless the reference itself
uses a tolerance."""
import math

import numpy as np
import pytest

F32 = np.float32


def eq(a, b):
    np.testing.assert_array_equal(np.asarray(a, dtype=F32), np.asarray(b, dtype=F32))


# ---- frames.rs -------------------------------------------------------------------------
def test_frames_from_slice(oracle):  # frames.rs:262-266
    f = oracle.Frames.from_slice(1, [1.0, 2.0, 3.0])
    assert f.len == 3
    s = oracle.FramesSignal(f, 0.0)
    eq(s.sample(1.0, 3), [1.0, 2.0, 3.0])


def test_frames_sample(oracle):  # frames.rs:269-275
    sig = oracle.FramesSignal(oracle.Frames.from_slice(1, [1.0, 2.0, 3.0, 4.0]), -2.0)
    eq(sig.sample(0.25, 4), [0.0, 0.0, 0.0, 0.0])
    eq(sig.sample(0.5, 3), [0.0, 0.5, 1.0])
    eq(sig.sample(1.0, 5), [1.5, 2.5, 3.5, 2.0, 0.0])


def test_frames_playback_position(oracle):  # frames.rs:278-303
    sig = oracle.FramesSignal(oracle.Frames.from_slice(1, [1.0, 2.0, 3.0]), -2.0)
    assert sig.playback_position() == -2.0
    assert not sig.control_is_finished()
    sig.sample(0.2, 10)
    assert sig.playback_position() == 0.0
    assert not sig.control_is_finished()
    sig.sample(0.1, 10)
    assert sig.playback_position() == 1.0
    sig.sample(0.1, 10)
    assert sig.playback_position() == 2.0
    sig.sample(0.2, 10)
    assert sig.control_is_finished()
    assert sig.playback_position() == 4.0
    sig.sample(0.5, 10)
    assert sig.playback_position() == 9.0


def test_frames_negative_fract_quirk(oracle):  # SURVEY Appendix A.2 (derived, frames.rs:189-195)
    sig = oracle.FramesSignal(oracle.Frames.from_slice(1, [1.0, 2.0, 3.0, 4.0]), -0.3)
    out = sig.sample(0.5, 4)
    np.testing.assert_allclose(out, [0.7, 1.2, 1.7, 2.2], rtol=0, atol=1e-6)


# ---- ring.rs -----------------------------------------------------------------------------
def test_ring_fill(oracle):  # ring.rs:106-120
    r = oracle.Ring(4)
    s = oracle.TimeSignal(1.0)
    r.write(s, 1, 1.0)
    assert r.write_cursor == 1.0
    eq(r.buffer, [1.0, 0.0, 0.0, 0.0])
    r.write(s, 1, 2.0)
    assert r.write_cursor == 3.0
    eq(r.buffer, [1.0, 2.0, 3.0, 0.0])
    eq(r.sample(1, -1.5, 1.0, 2), [2.5, 1.5])
    eq(r.sample(1, -1.5, 0.25, 4), [2.5, 2.75, 3.0, 2.25])


def test_ring_wrap(oracle):  # ring.rs:123-134
    r = oracle.Ring(4)
    s = oracle.TimeSignal(1.0)
    r.write(s, 1, 3.0)
    eq(r.buffer, [1.0, 2.0, 3.0, 0.0])
    r.write(s, 1, 3.0)
    eq(r.buffer, [5.0, 6.0, 3.0, 4.0])
    eq(r.sample(1, -2.75, 0.5, 6), [4.25, 4.75, 5.25, 5.75, 5.25, 3.75])


# ---- gain.rs / smooth.rs ------------------------------------------------------------------
def test_gain_smoothing(oracle):  # gain.rs:171-179
    g = oracle.Gain(oracle.Constant(1.0))
    g.control_set_amplitude_ratio(5.0)
    eq(g.sample(0.025, 6), [1.0, 2.0, 3.0, 4.0, 5.0, 5.0])
    eq(g.sample(0.025, 6), [5.0] * 6)


def test_smoothed_doctest(oracle):  # smooth.rs:7-24
    v = oracle.Smoothed(0.0)
    assert v.get() == 0.0
    v.set(1.0)
    assert v.get() == 0.0
    v.advance(0.5)
    assert v.get() == 0.5
    v.set(1.5)
    v.advance(0.5)
    assert v.get() == 1.0
    v.advance(0.5)
    assert v.get() == 1.5
    v.advance(0.5)
    assert v.get() == 1.5


# ---- mixer.rs ------------------------------------------------------------------------------
def test_mixer_is_stopped(oracle):  # mixer.rs:130-147
    mixer = oracle.Mixer(1)
    sig = oracle.FramesSignal(oracle.Frames.from_slice(1, [0.0, 0.0]), 0.0)
    handle = mixer.play(sig)
    assert not handle.is_stopped()
    mixer.sample(0.6, 1)
    assert not handle.is_stopped()
    mixer.sample(0.6, 1)
    assert not handle.is_stopped()  # finished, but not noticed until the next scan
    mixer.sample(0.0, 1)
    assert handle.is_stopped()


# ---- spatial.rs ------------------------------------------------------------------------------
def test_spatial_signal_finished(oracle):  # spatial.rs:631-665
    scene = oracle.SpatialScene()
    scene.play(oracle.FinishedSignal(), position=[343.0, 0.0, 0.0])
    scene.sample(0.0, 0)
    assert scene.len() == 1, "signal remains after no time has passed"
    scene.sample(0.6, 1)
    assert scene.len() == 1, "signal remains partway through propagation"
    scene.sample(0.6, 1)
    assert scene.len() == 1, "signal remains immediately after propagation delay expires"
    scene.sample(0.0, 0)
    assert scene.len() == 0, "signal dropped on first pass after propagation delay expires"


# ---- set.rs (membership only; capacities are a control-plane concern) ------------------------
def test_set_many_inserts(oracle):  # set.rs:227-251 realloc_signals / realloc_channel
    mixer = oracle.Mixer(2)
    frames = oracle.Frames.from_slice(10, np.zeros((10, 2), dtype=F32))
    for i in range(1, 131):
        mixer.play(oracle.FramesSignal(frames))
        mixer.sample(0.0, 0)
        assert len(mixer) == i
    mixer2 = oracle.Mixer(2)
    for _ in range(129):
        mixer2.play(oracle.FramesSignal(frames))
    assert len(mixer2) == 0
    mixer2.sample(0.0, 0)
    assert len(mixer2) == 129


# ---- math/mod.rs ------------------------------------------------------------------------------
def _axis_angle(axis, angle):  # math/mod.rs:131-142 -> mint layout [x, y, z, s]
    half = F32(angle) * F32(0.5)
    s, c = F32(math.sin(half)), F32(math.cos(half))
    return [axis[0] * s, axis[1] * s, axis[2] * s, c]


def test_rotate_x(oracle):  # math/mod.rs:102-109
    r = oracle.rotate(_axis_angle([1.0, 0.0, 0.0], math.pi / 2), [0.0, 0.0, -1.0])
    assert r[0] == 0.0 and abs(r[1] - 1.0) < 1e-3 and abs(r[2]) < 1e-3


def test_rotate_y(oracle):  # math/mod.rs:112-119
    r = oracle.rotate(_axis_angle([0.0, 1.0, 0.0], math.pi / 2), [1.0, 0.0, 0.0])
    assert abs(r[0]) < 1e-3 and r[1] == 0.0 and abs(r[2] + 1.0) < 1e-3


def test_rotate_z(oracle):  # math/mod.rs:122-129
    r = oracle.rotate(_axis_angle([0.0, 0.0, 1.0], math.pi / 2), [0.0, 1.0, 0.0])
    assert abs(r[1]) < 1e-3 and abs(r[0] + 1.0) < 1e-3 and r[2] == 0.0


# ---- signal.rs ---------------------------------------------------------------------------------
def test_mono_to_stereo(oracle):  # signal.rs:111-116
    s = oracle.MonoToStereo(oracle.CountingSignal(0))
    eq(s.sample(1.0, 4), [[0.0, 0.0], [1.0, 1.0], [2.0, 2.0], [3.0, 3.0]])


# ---- cycle.rs ("next" row; same gather with modulo) ------------------------------------------------
CYC = [1.0, 2.0, 3.0]


def test_cycle_wrap_single(oracle):  # cycle.rs:69-75
    s = oracle.Cycle(oracle.Frames.from_slice(1, CYC))
    eq(s.sample(1.0, 5), [1.0, 2.0, 3.0, 1.0, 2.0])


def test_cycle_wrap_multi(oracle):  # cycle.rs:77-84
    s = oracle.Cycle(oracle.Frames.from_slice(1, CYC))
    eq(np.concatenate([s.sample(1.0, 2), s.sample(1.0, 3)]), [1.0, 2.0, 3.0, 1.0, 2.0])


def test_cycle_wrap_fract(oracle):  # cycle.rs:86-93
    s = oracle.Cycle(oracle.Frames.from_slice(1, CYC))
    eq(np.concatenate([s.sample(0.5, 2), s.sample(0.5, 6)]), [1.0, 1.5, 2.0, 2.5, 3.0, 2.0, 1.0, 1.5])


def test_cycle_wrap_fract_offset(oracle):  # cycle.rs:95-103
    s = oracle.Cycle(oracle.Frames.from_slice(1, CYC))
    s.seek(0.25)
    eq(np.concatenate([s.sample(0.5, 2), s.sample(0.5, 5)]), [1.25, 1.75, 2.25, 2.75, 2.5, 1.5, 1.25])


def test_cycle_wrap_single_frame(oracle):  # cycle.rs:105-113
    s = oracle.Cycle(oracle.Frames.from_slice(1, [1.0]))
    s.seek(0.25)
    eq(np.concatenate([s.sample(1.0, 2), s.sample(1.0, 1)]), [1.0, 1.0, 1.0])


def test_cycle_wrap_large_interval(oracle):  # cycle.rs:115-122
    s = oracle.Cycle(oracle.Frames.from_slice(1, CYC))
    eq(np.concatenate([s.sample(10.0, 2), s.sample(10.0, 1)]), [1.0, 2.0, 3.0])


# ---- surveyor-derived vectors (NOT from the reference, cross-check only) ----------
APPENDIX_B = [
    ((1, 0, 0), (-0.003228863, 0.005599312), (-0.0026020408, 0.11691847)),
    ((0, 0, -1), (-0.0029322493, 0.06170182), (-0.0029322493, 0.06170182)),
    ((0, 0, 1), (-0.0029322493, 0.037725333), (-0.0029322493, 0.037725333)),
    ((-50, 10, 0), (-0.14835216, 0.0019193098), (-0.1489668, 4.957339e-05)),
    ((0, 0, 0), (-0.00031341109, 0.4651163), (-0.00031341109, 0.4651163)),
    ((343, 0, 0), (-1.0003135, 4.395334e-06), (-0.9996866, 0.00028732722)),
]


@pytest.mark.parametrize("p,left,right", APPENDIX_B)
def test_ear_state_appendix_b(oracle, p, left, right):
    for ear, want in ((0, left), (1, right)):
        off, gain = oracle.ear_state(p, ear, 0.1)
        assert off == pytest.approx(want[0], rel=2e-7, abs=0)
        assert gain == pytest.approx(want[1], rel=3e-6, abs=1e-12)


def test_rate_reciprocal_fast_path():  # SURVEY Appendix B note: f32(1/r)*f32(r) == 1.0 for common rates
    for r in (44100, 48000, 96000):
        assert F32(1.0) / F32(r) * F32(r) == F32(1.0)


def test_sine_phase_wrap_and_the_c1_scenario(oracle):
    """sine.rs:25-40 (no test in the reference: `Sine` is pinned here against an independent float32 statement of its
    three lines) and BASELINE.json's C1: 8 x MonoToStereo(Sine) under Mixer<[f32; 2]>, 1024-frame callbacks
    (examples/offline.rs:25-45 shape). The arguments of sinf are compared exactly; sinf itself is libm (glibc here, as
    for the Rust std build) and is only required to agree with numpy's float32 sine to 1 ulp."""
    o = oracle
    F = np.float32
    TAU = F(6.28318530717958647692)
    rate, n = 48000, 1024
    interval = F(1.0) / F(rate)
    rng = np.random.default_rng(18)
    params = [(F(rng.uniform(0, 2 * np.pi)), F(rng.uniform(100.0, 1000.0))) for _ in range(8)]
    sines = [o.Sine(float(ph), float(hz)) for ph, hz in params]
    mono_ref = []
    for (ph, hz), s in zip(params, sines):
        freq = F(hz * TAU)                                            # sine.rs:21
        phase = ph
        blocks = []
        for _ in range(3):                                            # three blocks: the phase wraps several times
            i = np.arange(n, dtype=F)
            arg = (interval * i).astype(F) * freq + phase             # :36-37, every operation rounded to f32
            arg = arg.astype(F)
            got = o.run(s, rate, n)
            want = np.sin(arg.astype(np.float64))
            assert np.max(np.abs(got.astype(np.float64) - want)) <= 1.2e-7   # <= 1 ulp at |x| <= 1
            blocks.append(got)
            phase = F(np.fmod(F(phase + F(F(interval * F(n)) * freq)), TAU))  # :25-28
            assert 0.0 <= phase < TAU
        mono_ref.append(blocks)
    # C1: the same eight sources (fresh state), duplicated to stereo and mixed in reverse set order (mixer.rs:100)
    mx = o.Mixer(2)
    for ph, hz in params:
        mx.play(o.MonoToStereo(o.Sine(float(ph), float(hz))))
    for b in range(3):
        out = o.run(mx, rate, n)
        acc = np.zeros(n, dtype=F)
        for k in reversed(range(8)):
            acc = (acc + mono_ref[k][b]).astype(F)
        np.testing.assert_array_equal(out[:, 0], acc)
        np.testing.assert_array_equal(out[:, 1], acc)                  # signal.rs:73-80 duplicates the sample
