"""Cycle (cycle.rs:6-61) on the device: plays under a Mixer through the literal kernel.
The reference's own six unit tests (cycle.rs:69-122), fed through `oddio::run`-style callbacks of a one-source
Mixer (0 + x == x, so the values are Cycle's own), then seeded parity against the oracle with the wrappers."""
import numpy as np
import pytest

from helpers import F32, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odb():
    import oddio_b200

    return oddio_b200


@pytest.fixture(scope="module")
def ctx(odb):
    return odb.init(0)


def play_cycle(odb, ctx, data, seek=None):
    fr = odb.Frames.from_slice(1, np.asarray(data, F32), ctx)
    cyc = odb.Cycle(fr)
    if seek is not None:
        cyc.seek(seek)
    ctl, mixer = odb.Mixer.new(1, ctx)
    ctl.play(cyc)
    return mixer


def chunks(mixer, interval, sizes):
    return np.concatenate([mixer.sample(interval, n) for n in sizes]).tolist()


FRAMES = [1.0, 2.0, 3.0]


def test_wrap_single(odb, ctx):  # cycle.rs:69-75
    assert chunks(play_cycle(odb, ctx, FRAMES), 1.0, [5]) == [1.0, 2.0, 3.0, 1.0, 2.0]


def test_wrap_multi(odb, ctx):  # cycle.rs:77-84
    assert chunks(play_cycle(odb, ctx, FRAMES), 1.0, [2, 3]) == [1.0, 2.0, 3.0, 1.0, 2.0]


def test_wrap_fract(odb, ctx):  # cycle.rs:86-93
    assert chunks(play_cycle(odb, ctx, FRAMES), 0.5, [2, 6]) == [1.0, 1.5, 2.0, 2.5, 3.0, 2.0, 1.0, 1.5]


def test_wrap_fract_offset(odb, ctx):  # cycle.rs:95-103
    assert chunks(play_cycle(odb, ctx, FRAMES, seek=0.25), 0.5, [2, 5]) == [1.25, 1.75, 2.25, 2.75, 2.5, 1.5, 1.25]


def test_wrap_single_frame(odb, ctx):  # cycle.rs:105-113
    assert chunks(play_cycle(odb, ctx, [1.0], seek=0.25), 1.0, [2, 1]) == [1.0, 1.0, 1.0]


def test_wrap_large_interval(odb, ctx):  # cycle.rs:115-122
    assert chunks(play_cycle(odb, ctx, FRAMES), 10.0, [2, 1]) == [1.0, 2.0, 3.0]


@pytest.mark.parametrize("channels", [1, 2])
def test_cycles_under_wrappers_match_the_oracle(oracle, odb, ctx, channels):
    """Many looping sources (short loops, so every callback wraps several times), bare and under Speed / FixedGain /
    Gain, next to ordinary FramesSignal sources; several callbacks of several mixer chunks; a gain change mid-run.
    Per source the values are bit-exact, so a one-source mixer is compared exactly and the sum within the protocol."""
    o = oracle
    rng = np.random.default_rng(60 + channels)
    rate = 48000
    ref, (ctl, dev) = o.Mixer(channels), odb.Mixer.new(channels, ctx)
    ref1, (ctl1, dev1) = o.Mixer(channels), odb.Mixer.new(channels, ctx)
    pcms = [synth_pcm(rng, int(n), rate, channels) for n in (37, 500, 1500, 4801)]
    long_pcm = synth_pcm(rng, 30000, rate, channels)
    cursors, gains = [], []
    for i in range(24):
        p = pcms[i % 4]
        fo, fd = o.Frames.from_slice(rate, p), odb.Frames.from_slice(rate, p, ctx)
        so, sd = o.Cycle(fo), odb.Cycle(fd)
        if i % 3 == 1:
            s = float(rng.uniform(-0.01, 0.01))
            so.seek(s); sd.seek(s)
        cursors.append((so, sd))
        io, idv = so, sd
        if i % 2:
            io = o.Speed(io); io.set_speed(float(F32(rng.uniform(0.5, 2.0))))
            sc, idv = odb.Speed.new(idv); sc.set_speed(io.speed() if hasattr(io, "speed") else 1.0)
        if i % 4 == 2:
            io, idv = o.FixedGain(io, -6.0), odb.FixedGain(idv, -6.0)
        if i % 5 == 0:
            io = o.Gain(io); io.set_amplitude_ratio(0.5)
            gc, idv = odb.Gain.new(idv); idv.set_amplitude_ratio(0.5)
            gains.append((io, gc))
        ref.play(io); ctl.play(idv)
    for _ in range(4):
        fo, fd = o.Frames.from_slice(rate, long_pcm), odb.Frames.from_slice(rate, long_pcm, ctx)
        ref.play(o.FramesSignal(fo, 0.0)); ctl.play(odb.FramesSignal(fd, 0.0))
    fo, fd = o.Frames.from_slice(rate, pcms[1]), odb.Frames.from_slice(rate, pcms[1], ctx)
    one_o, one_d = o.Cycle(fo), odb.Cycle(fd)
    ref1.play(one_o); ctl1.play(one_d)
    for k, n in enumerate((256, 1024, 3000, 4096, 33)):
        if k == 2:
            for io, gc in gains:
                io.control_set_amplitude_ratio(0.9); gc.set_amplitude_ratio(0.9)
        r = o.run(ref, rate, n)
        r64 = ref.out64(n)
        out = odb.run(dev, rate, np.zeros((n, channels) if channels > 1 else (n,), F32))
        rms = float(np.sqrt(np.mean(r64 ** 2)))
        tol = 1e-5 * np.maximum(np.abs(r.astype(np.float64)), rms)
        assert np.all(np.abs(out.astype(np.float64) - r) <= tol)
        np.testing.assert_array_equal(odb.run(dev1, rate, np.zeros_like(out)), o.run(ref1, rate, n))
        for so, sd in cursors:
            assert sd.control.cursor()[0] == so.cursor        # the f64 cursor, in samples (cycle.rs:52)
        assert one_d.control.cursor()[0] == one_o.cursor
    assert len(dev) == len(ref) == 28                          # a Cycle never finishes


def test_cycles_in_a_spatial_scene_match_the_oracle(oracle, odb, ctx):
    """SpatialSceneControl::play(Cycle) (Cycle is Seek, cycle.rs:56-61): looping point sources, moving and static,
    bare and under FixedGain, next to ordinary FramesSignal sources; the seeks of the mix closure (spatial.rs:449,
    :465, :468) go through Cycle's rem_euclid. One looping source alone is bit-exact; cursors always are."""
    from helpers import assert_mix_close, rand_in_shell

    o = oracle
    rng = np.random.default_rng(77)
    rate = 48000
    ref, (ctl, dev) = o.SpatialScene(), odb.SpatialScene.new(ctx)
    ref1, (ctl1, dev1) = o.SpatialScene(), odb.SpatialScene.new(ctx)
    pcms = [synth_pcm(rng, int(n), rate) for n in (211, 1999, 4801, 24000)]
    long_pcm = synth_pcm(rng, 60000, rate)
    cursors = []
    for i in range(20):
        p = pcms[i % 4]
        fo, fd = o.Frames.from_slice(rate, p), odb.Frames.from_slice(rate, p, ctx)
        so, sd = o.Cycle(fo), odb.Cycle(fd)
        cursors.append((so, sd))
        io, idv = (o.FixedGain(so, -3.0), odb.FixedGain(sd, -3.0)) if i % 4 == 3 else (so, sd)
        pos = rand_in_shell(rng, 2.0, 60.0)
        vel = rng.uniform(-30, 30, 3).astype(F32) if i % 2 else np.zeros(3, F32)
        ref.play(io, pos, vel, 0.1); ctl.play(idv, odb.SpatialOptions(pos, vel, 0.1))
    for _ in range(6):
        fo, fd = o.Frames.from_slice(rate, long_pcm), odb.Frames.from_slice(rate, long_pcm, ctx)
        pos, vel = rand_in_shell(rng, 2.0, 60.0), rng.uniform(-30, 30, 3).astype(F32)
        ref.play(o.FramesSignal(fo, 1.0), pos, vel, 0.1); ctl.play(odb.FramesSignal(fd, 1.0), odb.SpatialOptions(pos, vel, 0.1))
    fo, fd = o.Frames.from_slice(rate, pcms[1]), odb.Frames.from_slice(rate, pcms[1], ctx)
    one_o, one_d = o.Cycle(fo), odb.Cycle(fd)
    ref1.play(one_o, [3.0, 1.0, -2.0], [10.0, -3.0, 4.0], 0.1)
    ctl1.play(one_d, odb.SpatialOptions([3.0, 1.0, -2.0], [10.0, -3.0, 4.0], 0.1))
    for n in (256, 1024, 1500, 33, 4096):
        r = o.run(ref, rate, n)
        out = odb.run(dev, rate, np.zeros((n, 2), F32))
        assert_mix_close(out, r, ref.out64(n))
        np.testing.assert_array_equal(odb.run(dev1, rate, np.zeros((n, 2), F32)), o.run(ref1, rate, n))
        for so, sd in cursors:
            assert sd.control.cursor()[0] == so.cursor
        assert one_d.control.cursor()[0] == one_o.cursor
    assert dev.len() == ref.len() == 26
    cnt = dev.last_job_counters()
    assert cnt["general"] == 20 * 4 and cnt["staged"] == 6 * 4   # 4096 frames = 4 tiles: loops literal, the rest staged


def test_cycle_is_rejected_by_play_buffered(odb, ctx):
    fr = odb.Frames.from_slice(48000, np.zeros(100, F32), ctx)
    ctl, scene = odb.SpatialScene.new(ctx)
    with pytest.raises(odb.OddioError):
        ctl.play_buffered(odb.Cycle(fr), odb.SpatialOptions([1.0, 0.0, 0.0], [0.0, 0.0, 0.0], 0.1), 100.0, 48000, 0.1)
