"""Parity of the device Mixer (mixer.rs) over Gain / FixedGain / Speed / FramesSignal chains, with the
Tanh / Reinhard wrappers, against the CPU oracle, through the C ABI.

Kernel variants: 0 = streaming kernel for ds ~= 1 sources + literal kernel for the rest (default);
1 = literal kernel for every source. Single-source outputs are bit-exact (no summation freedom)."""
import numpy as np
import pytest

from helpers import F32, MixerPair, assert_mix_close, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odb():
    import oddio_b200

    return oddio_b200


@pytest.fixture(scope="module")
def ctx(odb):
    return odb.init(0)


def check_cursors(pair):
    for it in pair.items:
        try:
            t, _ = it["dev_frames_control"].cursor()
        except Exception:
            continue
        assert t == it["ref_frames_signal"].t
        assert it["dev_frames_control"].playback_position() == it["ref_frames_signal"].playback_position()


def close(out, ref, ref64, epilogue):
    if epilogue is None:
        assert_mix_close(out, ref, ref64.reshape(ref.shape))
    else:  # the f64 aid holds the pre-limiter mix; compare the limited outputs directly
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("channels", [1, 2])
def test_single_source_bit_exact(oracle, odb, ctx, variant, channels):
    rng = np.random.default_rng(10 + channels)
    pcm = synth_pcm(rng, 20000, 48000, channels)
    for kw in ({}, {"gain": 0.37}, {"fixed_gain_db": -4.5}, {"speed": 1.37}, {"speed": 0.61, "gain": 0.8, "fixed_gain_db": 2.0}):
        pair = MixerPair(oracle, odb, ctx, channels)
        pair.dev_mixer.set_kernel_variant(variant)
        pair.play(48000, pcm, 0.0, **kw)
        for n in (256, 1024, 1500, 4096, 7):
            ref, _, out = pair.step(48000, n)
            np.testing.assert_array_equal(out, ref)
            check_cursors(pair)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("epilogue", [None, "tanh", "reinhard"])
def test_c4_like_static_stereo_gain_tanh(oracle, odb, ctx, variant, epilogue):
    """BASELINE.json configs[3] in small: static stereo FramesSignal sources under Gain, whole mixer under Tanh, 96 kHz."""
    rng = np.random.default_rng(20)
    rate = 96000
    pair = MixerPair(oracle, odb, ctx, 2, epilogue)
    pair.dev_mixer.set_kernel_variant(variant)
    pcms = [synth_pcm(rng, 9000, rate, 2) for _ in range(8)]
    n_src = 300
    for i in range(n_src):
        pair.play(rate, pcms[i % 8], 0.0, gain=float(rng.uniform(0.05, 1.0)) * (0.02 if epilogue is None else 0.05))
    for _ in range(4):
        ref, ref64, out = pair.step(rate, 1024)
        close(out, ref, ref64, epilogue)
        check_cursors(pair)
    cnt = pair.dev_mixer.last_job_counters()
    assert cnt == ({"general": 0, "staged": n_src, "resampled": 0, "ring_literal": 0} if variant == 0 else {"general": n_src, "staged": 0, "resampled": 0, "ring_literal": 0})


@pytest.mark.parametrize("variant", [0, 1])
def test_c5_like_speed_sweep(oracle, odb, ctx, variant):
    """BASELINE.json configs[4] in small: mono Speed<FramesSignal> sources, ratio 0.5..2.0, 4096-frame buffer."""
    rng = np.random.default_rng(30)
    rate = 48000
    pair = MixerPair(oracle, odb, ctx, 1)
    pair.dev_mixer.set_kernel_variant(variant)
    pcms = [synth_pcm(rng, 30000, rate, 1) for _ in range(4)]
    for i in range(64):
        pair.play(rate, pcms[i % 4], 0.0, speed=float(rng.uniform(0.5, 2.0)))
    for _ in range(3):
        ref, ref64, out = pair.step(rate, 4096)
        close(out, ref, ref64, None)
        check_cursors(pair)
    cnt = pair.dev_mixer.last_job_counters()  # 64 sources x 4 chunks: all on the staged resampling kernel
    assert cnt == ({"general": 0, "staged": 0, "resampled": 256, "ring_literal": 0} if variant == 0 else {"general": 256, "staged": 0, "resampled": 0, "ring_literal": 0})


@pytest.mark.parametrize("variant", [0, 1])
def test_gain_transitions_speed_changes_stop_and_finish(oracle, odb, ctx, variant):
    """Gain's 0.1 s smoothing ramps (gain.rs:103-122), set_speed mid-run, Mixed::stop,
    sources running out and being dropped one callback late (mixer.rs:102-106)."""
    rng = np.random.default_rng(40)
    rate = 48000
    pair = MixerPair(oracle, odb, ctx, 2)
    pair.dev_mixer.set_kernel_variant(variant)
    long_pcm = synth_pcm(rng, 40000, rate, 2)
    short_pcm = synth_pcm(rng, 3000, rate, 2)
    for i in range(12):
        pair.play(rate, long_pcm, 0.01 * i, speed=1.0 if i % 2 else None, gain=0.5)
    for i in range(4):
        pair.play(rate, short_pcm, 0.0, gain=0.9)
    for step in range(14):
        if step == 1:
            for i in range(0, 12, 3):
                pair.set_gain_ratio(i, float(rng.uniform(0.1, 2.0)))
        if step == 3:
            pair.set_gain_ratio(0, 1.0)  # re-target in the middle of a ramp
            pair.set_speed(1, 1.25)
            pair.set_speed(3, 0.75)
        if step == 5:
            pair.stop(2)
            assert pair.items[2]["dev_mixed"].is_stopped() and pair.items[2]["ref_mixed"].is_stopped()
        ref, ref64, out = pair.step(rate, 1024 if step % 3 else 640)
        close(out, ref, ref64, None)
        check_cursors(pair)
        assert len(pair.dev_mixer) == len(pair.ref_mixer)
        for it in pair.items:
            assert it["dev_mixed"].is_stopped() == it["ref_mixed"].is_stopped()
    assert len(pair.ref_mixer) == 11


def test_reference_is_stopped_test_on_device(odb, ctx):
    """mixer.rs:130-147 replayed through the C ABI: a 2-frame source @1 Hz is dropped one callback late."""
    ctl, mixer = odb.Mixer.new(1, ctx)
    frames = odb.Frames.from_slice(1, np.array([0.0, 0.0], dtype=F32), ctx)
    handle = ctl.play(odb.FramesSignal(frames, 0.0))
    mixer.sample(0.6, 1)
    assert not handle.is_stopped()
    mixer.sample(0.6, 1)
    assert not handle.is_stopped()
    mixer.sample(0.0, 1)
    assert handle.is_stopped()


def test_rate_mismatch_and_negative_start(oracle, odb, ctx):
    rng = np.random.default_rng(50)
    pair = MixerPair(oracle, odb, ctx, 1)
    for rate in (22050, 44100, 96000):
        pcm = synth_pcm(rng, rate // 2, rate, 1)
        pair.play(rate, pcm, -0.004)
        pair.play(rate, pcm, 0.1, gain=0.3)
    for n in (512, 2048):
        ref, ref64, out = pair.step(48000, n)
        close(out, ref, ref64, None)
        check_cursors(pair)


def test_empty_mixer(oracle, odb, ctx):
    pair = MixerPair(oracle, odb, ctx, 2)
    ref, _, out = pair.step(48000, 128)
    np.testing.assert_array_equal(out, ref)


def test_callback_of_48000_frames(oracle, odb, ctx):
    """mixer.rs:109-117 chunks any out.len() through its 1024-frame staging buffer."""
    rng = np.random.default_rng(4800)
    rate, n = 48000, 48000
    pair = MixerPair(oracle, odb, ctx, 2)
    pcms = [synth_pcm(rng, 2 * n * 2 + 4096, rate, 2) for _ in range(3)]
    for i in range(9):
        pair.play(rate, pcms[i % 3], 0.01 * i, speed=float(rng.uniform(0.7, 1.6)) if i % 3 else None,
                  gain=float(rng.uniform(0.1, 1.0)))
    for _ in range(2):
        ref, ref64, out = pair.step(rate, n)
        assert_mix_close(out, ref, ref64)
    for it in pair.items:
        assert it["dev_frames_control"].cursor()[0] == it["ref_frames_signal"].t
