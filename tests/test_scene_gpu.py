"""Parity of the device SpatialScene (seek path) against the CPU oracle, through the C ABI.

Kernel variants (odb_set_kernel_variant): 0 = staged mix, strict arithmetic; 1 = the literal path for every
source; 2 = staged mix with FMA-contracted value operations (the library's default); + 0x100 = per-source
set-up kernels on a second stream (overlapping the previous callback's mix); + 0x200 = round 1's multi-kernel
callback (walk, k_mix_fast, k_mix_general, k_reduce_tiles) instead of the one-launch kernel.

Bars: f64 time cursors bit-exact; a single source's contribution bit-exact in the
strict build; mixed output within 1e-5 * max(|ref|, RMS)."""
import numpy as np
import pytest

from helpers import F32, ScenePair, assert_mix_close, rand_in_shell, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odb():
    import oddio_b200

    return oddio_b200


@pytest.fixture(scope="module")
def ctx(odb):
    return odb.init(0)


def cursors_equal(pair):
    for so, cd in zip(pair.ref_signals, pair.dev_controls):
        try:
            t, _ = cd.cursor()
        except Exception:
            continue  # removed on the device; the oracle object outlives its removal
        assert t == so.t, f"f64 cursor differs: {t!r} vs {so.t!r}"


@pytest.mark.parametrize("variant", [0, 1, 2, 0x200, 0x202])
def test_single_source_bit_exact(oracle, odb, ctx, variant):
    """One moving source: no summation-order freedom, so the output must equal the oracle bit for bit."""
    rng = np.random.default_rng(1)
    rate = 48000
    pcm = synth_pcm(rng, 60000, rate)
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    pair.play(rate, pcm, 0.5, [3.0, 1.0, -2.0], [10.0, -3.0, 4.0])
    for n in (256, 1024, 100, 1, 777, 2048, 4096, 3000):
        ref, ref64, out = pair.step(rate, n)
        if variant & 0xFF == 2:  # value multiply-adds contracted to FMA: <= 1e-5 relative, indices still exact
            assert_mix_close(out, ref, ref64)
        else:
            np.testing.assert_array_equal(out, ref)
        cursors_equal(pair)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_static_source_fast_path_bit_exact(oracle, odb, ctx, variant):
    """Static source at the output rate: FramesSignal's ds ~= 1 fast path (frames.rs:180-187)."""
    rng = np.random.default_rng(2)
    rate = 48000
    pcm = synth_pcm(rng, 30000, rate)
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    pair.play(rate, pcm, 0.25, [0.0, 0.0, -5.0], [0.0, 0.0, 0.0])
    for n in (256, 512, 1024, 33):
        ref, ref64, out = pair.step(rate, n)
        if variant & 0xFF == 2:
            assert_mix_close(out, ref, ref64)
        else:
            np.testing.assert_array_equal(out, ref)
        cursors_equal(pair)


@pytest.mark.parametrize("variant", [0, 1, 2, 0x100, 0x102, 0x200, 0x201, 0x202, 0x300])
@pytest.mark.parametrize("n_src,n_frames", [(8, 256), (300, 256), (1024, 256), (515, 1024), (64, 2048), (33, 1500)])
def test_many_sources(oracle, odb, ctx, variant, n_src, n_frames):
    rng = np.random.default_rng(100 + n_src)
    rate = 48000
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    pcms = [synth_pcm(rng, 48000 + 6 * n_frames * 2, rate) for _ in range(min(n_src, 16))]
    for i in range(n_src):
        pos = rand_in_shell(rng, 2.0, 100.0)
        vel = rng.uniform(-30, 30, 3).astype(F32)
        pair.play(rate, pcms[i % len(pcms)], 1.0, pos, vel)
    for _ in range(4):
        ref, ref64, out = pair.step(rate, n_frames)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)
    assert pair.dev.len() == pair.ref.len() == n_src
    cnt = pair.dev.last_job_counters()
    tiles = (n_frames + 1023) // 1024
    if variant & 0xFF == 1:
        assert cnt == {"general": n_src * tiles, "staged": 0, "resampled": 0, "ring_literal": 0}
    else:  # sources well inside their PCM, |ds - 1| < 0.4: the staged kernel must take all of them
        assert cnt == {"general": 0, "staged": n_src * tiles, "resampled": 0, "ring_literal": 0}


@pytest.mark.parametrize("variant", [0, 1, 2, 0x200, 0x202])
def test_start_before_zero_and_run_off_the_end(oracle, odb, ctx, variant):
    """Sources that start at negative time (zeros, then the k == -1 pair, frames.rs:118-122) and
    sources that run past their last frame and get dropped once the tail has propagated
    (spatial.rs:243-261)."""
    rng = np.random.default_rng(3)
    rate = 48000
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    short = synth_pcm(rng, 1500, rate)
    pair.play(rate, short, -0.01, [1.0, 0.5, 0.0], [0.0, 0.0, 0.0])
    pair.play(rate, short, 0.0, [20.0, 0.0, 3.0], [5.0, 0.0, 0.0])
    pair.play(rate, short, 0.02, [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    pair.play(rate, short, -0.3, [0.3, 0.0, 0.0], [0.0, 0.0, 0.0])
    lens = []
    for _ in range(40):
        ref, ref64, out = pair.step(rate, 256)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)
        assert pair.dev.len() == pair.ref.len()
        lens.append(pair.ref.len())
        for hr, hd in zip(pair.ref_handles, pair.dev_handles):
            assert hr.is_finished() == hd.is_finished()
    assert lens[0] == 4 and lens[-1] < 4


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_motion_updates_and_listener_rotation(oracle, odb, ctx, variant):
    rng = np.random.default_rng(4)
    rate = 48000
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    pcms = [synth_pcm(rng, 80000, rate) for _ in range(4)]
    n_src = 50
    for i in range(n_src):
        pair.play(rate, pcms[i % 4], 1.0, rand_in_shell(rng, 2, 50), rng.uniform(-10, 10, 3).astype(F32), radius=0.5)
    for step in range(8):
        for i in rng.choice(n_src, 7, replace=False):
            pair.set_motion(int(i), rand_in_shell(rng, 2, 50), rng.uniform(-10, 10, 3).astype(F32), bool(rng.integers(2)))
        if step % 2 == 1:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            pair.set_listener_rotation(q.astype(F32))
        ref, ref64, out = pair.step(rate, 512)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_resampled_pcm_rates_and_fixed_gain(oracle, odb, ctx, variant):
    """PCM at rates other than the output rate (ds far from 1: 0.46 ... 2.0) and FixedGain (gain.rs:32-37)."""
    rng = np.random.default_rng(5)
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    for rate in (22050, 44100, 48000, 96000):
        pcm = synth_pcm(rng, rate * 2, rate)
        for _ in range(5):
            pair.play(rate, pcm, 0.8, rand_in_shell(rng, 1, 30), rng.uniform(-20, 20, 3).astype(F32),
                      fixed_gain_db=float(rng.uniform(-12, 6)) if rng.integers(2) else None)
    for n in (256, 1024, 333):
        ref, ref64, out = pair.step(48000, n)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)


def test_empty_scene_and_zero_frames(oracle, odb, ctx):
    pair = ScenePair(oracle, odb, ctx)
    ref, _, out = pair.step(48000, 256)
    np.testing.assert_array_equal(out, ref)
    rng = np.random.default_rng(6)
    pair.play(48000, synth_pcm(rng, 5000, 48000), 0.0, [1, 2, 3], [0, 0, 0])
    ref, _, out = pair.step(48000, 0)
    assert out.shape == (0, 2)
    ref, _, out = pair.step(48000, 64)
    np.testing.assert_array_equal(out, ref)


def test_playback_position_readback(oracle, odb, ctx):
    rng = np.random.default_rng(7)
    pair = ScenePair(oracle, odb, ctx)
    i = pair.play(48000, synth_pcm(rng, 48000, 48000), 0.1, [4, 0, 0], [1, 0, 0])
    assert pair.dev_controls[i].playback_position() == pair.ref_signals[i].playback_position()
    for _ in range(3):
        pair.step(48000, 480)
        assert pair.dev_controls[i].playback_position() == pair.ref_signals[i].playback_position()
        assert pair.dev_controls[i].is_finished() == pair.ref_signals[i].control_is_finished()


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("n_frames", [8192, 48000, 5000])
def test_callbacks_longer_than_four_tiles(oracle, odb, ctx, variant, n_frames):
    """spatial.rs:456 takes any out.len(): the walk and the one-launch kernel loop over the 1024-frame tiles."""
    rng = np.random.default_rng(8000 + n_frames)
    rate = 48000
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)
    pcms = [synth_pcm(rng, 48000 + int(1.3 * n_frames * 2) + 4096, rate) for _ in range(4)]
    for i in range(19):
        pair.play(rate, pcms[i % 4], 1.0, rand_in_shell(rng, 2.0, 100.0), rng.uniform(-30, 30, 3).astype(F32))
    pair.play(rate, pcms[0], 1.0, [0.0, 0.0, -3.0], [0.0, 0.0, 0.0])   # static: the ds ~= 1 path
    pair.play(rate, pcms[1], 0.5, rand_in_shell(rng, 2.0, 50.0), rng.uniform(-30, 30, 3).astype(F32), fixed_gain_db=-6.0)  # literal path
    for _ in range(2):
        ref, ref64, out = pair.step(rate, n_frames)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)
    one = ScenePair(oracle, odb, ctx)   # a single moving source: bit for bit in the strict build
    one.dev.set_kernel_variant(variant)
    one.play(rate, pcms[2], 1.0, [30.0, 5.0, -20.0], [-20.0, 3.0, 11.0])
    ref, ref64, out = one.step(rate, n_frames)
    if variant == 0:
        np.testing.assert_array_equal(out, ref)
    else:
        assert_mix_close(out, ref, ref64)
