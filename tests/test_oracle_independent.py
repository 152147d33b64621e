"""The C++ oracle against the second, independent restatement of the hot path (tests/independent.py: scalar numpy
float32 / float64 operations written from the reference alone): bit-identical outputs, f64 cursors and set membership.
See the module docstring of tests/independent.py for what is restated and what is not (libm)."""
import numpy as np
import pytest

from independent import PyFixedGain, PyFramesSignal, PyGain, PyMixer, PyScene, PySpeed, f32, make_pcm


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scene_sample_seek_path_agrees_bit_for_bit(oracle, seed):
    rng = np.random.default_rng(seed)
    rate = 48000
    ref, py = oracle.SpatialScene(), PyScene()
    ref_src, py_src, ref_sig, py_sig = [], [], [], []
    for i in range(4):
        pcm_rate = rate if i != 2 else 44100
        pcm = make_pcm(rng, 3000 if i == 3 else 9000, pcm_rate)  # source 3 runs off its end inside the test
        start = float(rng.uniform(-0.01, 0.05))
        d = rng.normal(size=3)
        pos = (d / np.linalg.norm(d) * rng.uniform(1.0, 60.0)).astype(f32)
        vel = np.zeros(3, f32) if i == 1 else rng.uniform(-40, 40, 3).astype(f32)  # source 1: the ds ~= 1 path
        radius = float(rng.uniform(0.05, 0.5))
        sig = oracle.FramesSignal(oracle.Frames.from_slice(pcm_rate, pcm), start)
        ref_sig.append(sig)
        ref_src.append(ref.play(sig, pos, vel, radius))
        inner = PyFramesSignal(pcm, pcm_rate, start)
        py_sig.append(inner)
        py_src.append(py.play(inner, pos, vel, radius))
    for step, n in enumerate((256, 300, 1, 1024, 700, 513, 2048, 4096, 4096, 4096, 1024)):
        if step == 2:  # a motion update with and without a discontinuity, and a listener turn
            for j, disc in ((0, False), (2, True)):
                pos, vel = rng.uniform(-30, 30, 3).astype(f32), rng.uniform(-20, 20, 3).astype(f32)
                ref_src[j].set_motion(pos, vel, disc)
                py_src[j].set_motion(pos, vel, disc)
            q = rng.normal(size=4)
            q = (q / np.linalg.norm(q)).astype(f32)
            ref.set_listener_rotation(q)
            py.set_listener_rotation(q)
        a = oracle.run(ref, rate, n)
        b = py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"callback {step} ({n} frames)")
        assert ref.len() == len(py.sources)
        live = {id(s.inner) for s in py.sources}
        for rs, ps in zip(ref_sig, py_sig):
            if id(ps) in live:
                assert rs.t == float(ps.t)
    assert len(py.sources) < 4, "the short source should have been dropped after its propagation delay"


@pytest.mark.parametrize("channels", [1, 2])
def test_mixer_chain_agrees_bit_for_bit(oracle, channels):
    rng = np.random.default_rng(10 + channels)
    rate = 48000
    ref, py = oracle.Mixer(channels), PyMixer(channels)
    items = []
    for i in range(6):
        pcm_rate = 44100 if i == 4 else rate
        n = 2500 if i == 5 else 30000  # signal 5 finishes inside the test
        pcm = np.stack([make_pcm(rng, n, pcm_rate) for _ in range(channels)], axis=1) if channels > 1 else make_pcm(rng, n, pcm_rate)
        start = float(rng.uniform(0.0, 0.02))
        o = oracle.FramesSignal(oracle.Frames.from_slice(pcm_rate, pcm), start)
        p = PyFramesSignal(pcm, pcm_rate, start)
        it = {"ref_frames": o, "py_frames": p}
        if i in (1, 3):
            o, p = oracle.Speed(o), PySpeed(p)
            it["ref_speed"], it["py_speed"] = o, p
            o.set_speed(0.7)
            p.speed = f32(0.7)
        if i in (2, 3):
            o = oracle.FixedGain(o, -6.0)
            p = PyFixedGain(p, o.gain)
        if i in (0, 3):
            o, p = oracle.Gain(o), PyGain(p)
            it["ref_gain"], it["py_gain"] = o, p
            if i == 3:
                o.set_amplitude_ratio(0.5)
                p.set_amplitude_ratio(0.5)
        it["ref_mixed"], it["py_mixed"] = ref.play(o), py.play(p)
        items.append(it)
    for step, n in enumerate((256, 1024, 1500, 3000, 700, 4096, 2048)):
        if step == 2:  # a gain transition (0.1 s = 4800 frames: it spans callbacks and the 1024-frame staging chunks)
            items[0]["ref_gain"].control_set_amplitude_ratio(0.25)
            items[0]["py_gain"].control_set_amplitude_ratio(0.25)
            items[1]["ref_speed"].set_speed(1.3)
            items[1]["py_speed"].speed = f32(1.3)
        if step == 3:  # retargeted in mid-transition
            items[0]["ref_gain"].control_set_amplitude_ratio(0.8)
            items[0]["py_gain"].control_set_amplitude_ratio(0.8)
        if step == 4:
            items[2]["ref_mixed"].stop()
            items[2]["py_mixed"]["stop"] = True
        a, b = oracle.run(ref, rate, n), py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"callback {step} ({n} frames)")
        assert len(ref) == len(py.signals)
        for it in items:
            if any(e is it["py_mixed"] for e in py.signals):
                assert it["ref_frames"].t == float(it["py_frames"].t)
    assert len(py.signals) == 4, "one signal stopped, one ran off its end"


@pytest.mark.parametrize("seed", [0, 1])
def test_scene_sample_buffered_path_agrees_bit_for_bit(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    rate = 48000
    ref, py = oracle.SpatialScene(), PyScene()
    ref_src, py_src = [], []
    for i in range(3):
        pcm_rate = 44100 if i == 1 else rate
        pcm = make_pcm(rng, 4000 if i == 2 else 20000, pcm_rate)  # source 2 runs off its end
        start = float(rng.uniform(0.0, 0.03))
        d = rng.normal(size=3)
        pos = (d / np.linalg.norm(d) * rng.uniform(1.0, 40.0)).astype(f32)
        vel = rng.uniform(-30, 30, 3).astype(f32)
        radius = float(rng.uniform(0.05, 0.5))
        ring_rate = 48000 if i != 1 else 32000
        o = oracle.FramesSignal(oracle.Frames.from_slice(pcm_rate, pcm), start)
        p = PyFramesSignal(pcm, pcm_rate, start)
        if i == 0:  # a chain that does not implement Seek: what play_buffered is for
            o, p = oracle.Speed(o), PySpeed(p)
            o.set_speed(1.1)
            p.speed = f32(1.1)
        ref_src.append(ref.play_buffered(o, pos, vel, radius, 50.0, ring_rate, 0.1))
        py_src.append(py.play_buffered(p, pos, vel, radius, 50.0, ring_rate, 0.1))
    for step, n in enumerate((256, 1024, 300, 2048, 4096, 4096, 4096, 1000)):
        if step == 2:
            pos, vel = rng.uniform(-20, 20, 3).astype(f32), rng.uniform(-20, 20, 3).astype(f32)
            ref_src[1].set_motion(pos, vel, False)
            py_src[1].set_motion(pos, vel, False)
            q = rng.normal(size=4)
            q = (q / np.linalg.norm(q)).astype(f32)
            ref.set_listener_rotation(q)
            py.set_listener_rotation(q)
        a, b = oracle.run(ref, rate, n), py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"callback {step} ({n} frames)")
        assert ref.len(buffered=True) == len(py.buffered)
    assert len(py.buffered) == 2, "the short source should have been dropped after its propagation delay"


def test_both_sets_and_edge_cases_agree_bit_for_bit(oracle):
    """Both sets in one scene (one output buffer: the buffered set is mixed first, spatial.rs:395, 435), with the corner
    cases: a source sitting exactly in an ear (distance < 1e-3, spatial.rs:528-530), one that starts before its first
    sample (get_pair's negative indices, frames.rs:118-122), one passing through the centre of the head, a buffered
    source beyond max_distance (the clamp at -max_delay, spatial.rs:411-412), one-frame and 257-frame callbacks."""
    rng = np.random.default_rng(7)
    rate = 48000
    ref, py = oracle.SpatialScene(), PyScene()
    pcm = make_pcm(rng, 12000, rate)
    cases = [
        ("seek", [-0.1075, 0.0, 0.0], [0.0, 0.0, 0.0], 0.0),      # in the left ear, static
        ("seek", [4.0, -2.0, 1.0], [3.0, 1.0, -8.0], -0.05),      # starts 2400 frames before its first sample
        ("seek", [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 0.01),         # through the centre of the head
        ("buffered", [60.0, 5.0, -3.0], [-5.0, 0.0, 2.0], 0.0),   # beyond max_distance = 50 m
        ("buffered", [0.5, 0.2, -0.1], [0.0, 0.0, 0.0], 0.002),   # close and static
    ]
    for kind, pos, vel, start in cases:
        o = oracle.FramesSignal(oracle.Frames.from_slice(rate, pcm), start)
        p = PyFramesSignal(pcm, rate, start)
        pos, vel = np.array(pos, f32), np.array(vel, f32)
        if kind == "seek":
            ref.play(o, pos, vel, 0.1)
            py.play(p, pos, vel, 0.1)
        else:
            ref.play_buffered(o, pos, vel, 0.1, 50.0, rate, 0.1)
            py.play_buffered(p, pos, vel, 0.1, 50.0, rate, 0.1)
    for n in (1, 257, 1024, 300, 2048):
        a, b = oracle.run(ref, rate, n), py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"{n} frames")
        assert ref.len() == len(py.sources) and ref.len(buffered=True) == len(py.buffered)


def test_c1_sines_and_reinhard_agree_bit_for_bit(oracle):
    """BASELINE config 1 (examples/offline.rs: Sine -> MonoToStereo -> Mixer) and Reinhard over the mixer; sinf is
    glibc's on both sides, the phase wrap (sine.rs:25-28) and everything around it are restated."""
    from independent import PyMonoToStereo, PySine, reinhard32

    rng = np.random.default_rng(21)
    rate = 48000
    ref_mx, py = oracle.Mixer(2), PyMixer(2)
    ref = oracle.Reinhard(ref_mx)
    for _ in range(8):
        phase, freq = float(f32(rng.uniform(0, 6.28))), float(f32(rng.uniform(100.0, 1000.0)))
        ref_mx.play(oracle.MonoToStereo(oracle.Sine(phase, freq)))
        py.play(PyMonoToStereo(PySine(phase, freq)))
    for n in (1024, 1024, 300, 2048):
        a, b = oracle.run(ref, rate, n), reinhard32(py.run(rate, n))
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"{n} frames")


@pytest.mark.parametrize("channels", [1, 2])
def test_cycle_under_the_mixer_agrees_bit_for_bit(oracle, channels):
    from independent import PyCycle

    rng = np.random.default_rng(30 + channels)
    rate = 48000
    ref, py = oracle.Mixer(channels), PyMixer(channels)
    pairs = []
    for i in range(3):
        pcm_rate = 44100 if i == 1 else rate
        n = (700, 1531, 64)[i]  # short loops: many wraps per callback, one shorter than a callback's advance
        pcm = np.stack([make_pcm(rng, n, pcm_rate) for _ in range(channels)], axis=1) if channels > 1 else make_pcm(rng, n, pcm_rate)
        o, p = oracle.Cycle(oracle.Frames.from_slice(pcm_rate, pcm)), PyCycle(pcm, pcm_rate)
        pairs.append((o, p))
        if i == 2:
            o, p = oracle.Speed(o), PySpeed(p)
            o.set_speed(1.7)
            p.speed = f32(1.7)
        ref.play(o)
        py.play(p)
    for n in (256, 1024, 1500, 4096):
        a, b = oracle.run(ref, rate, n), py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"{n} frames")
        for o, p in pairs:
            assert o.cursor == float(p.cursor)


def test_cycle_in_the_scene_agrees_bit_for_bit(oracle):
    """Cycle is Seek (cycle.rs:56-61): under SpatialSceneControl::play its cursor goes through the five seeks per
    callback and wraps with rem_euclid."""
    from independent import PyCycle

    rng = np.random.default_rng(40)
    rate = 48000
    ref, py = oracle.SpatialScene(), PyScene()
    pairs = []
    for i in range(2):
        pcm = make_pcm(rng, (900, 5000)[i], rate)
        o, p = oracle.Cycle(oracle.Frames.from_slice(rate, pcm)), PyCycle(pcm, rate)
        pairs.append((o, p))
        pos, vel = rng.uniform(-20, 20, 3).astype(f32), rng.uniform(-20, 20, 3).astype(f32)
        ref.play(o, pos, vel, 0.1)
        py.play(p, pos, vel, 0.1)
    for n in (256, 1024, 700, 2048):
        a, b = oracle.run(ref, rate, n), py.run(rate, n)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"{n} frames")
        for o, p in pairs:
            assert o.cursor == float(p.cursor)
