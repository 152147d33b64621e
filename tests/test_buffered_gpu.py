"""Parity of buffered spatial sources (`play_buffered`: per-source delay ring in HBM, ring.rs +
spatial.rs:395-433) against the CPU oracle, through the C ABI. The ring's f32 write cursor and the inner
FramesSignal's f64 cursor must be bit-exact; a single source's output is bit-exact."""
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
            continue
        assert t == so.t, f"f64 cursor differs: {t!r} vs {so.t!r}"


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("kw", [{}, {"gain": 0.4}, {"fixed_gain_db": -3.0}, {"speed": 1.3}, {"speed": 0.7, "gain": 1.7}])
def test_single_buffered_source_bit_exact(oracle, odb, ctx, kw, variant):
    rng = np.random.default_rng(60)
    rate = 48000
    pcm = synth_pcm(rng, 140000, rate)
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(variant)  # 0: staged kernel reads the ring (literal ring kernel when a read wraps); 1: literal only
    pair.play_buffered(rate, pcm, 0.0, [30.0, 5.0, -12.0], [12.0, -3.0, 4.0], 0.2, max_distance=200.0, ring_rate=rate,
                       buffer_duration=0.1, **kw)
    kinds = set()
    for n in (256, 1024, 480, 1024, 2048, 100) * 6:  # ~0.6 s: the 0.68 s ring wraps along the way
        ref, _, out = pair.step(rate, n)
        np.testing.assert_array_equal(out, ref)
        cursors_equal(pair)
        cnt = pair.dev.last_job_counters()
        kinds.add("staged" if cnt["staged"] else "literal")
    assert np.abs(out).max() > 0  # the sound has arrived (30 m ~ 87 ms)
    assert kinds == ({"staged", "literal"} if variant == 0 else {"literal"})


def test_many_buffered_and_seek_sources_together(oracle, odb, ctx):
    rng = np.random.default_rng(61)
    rate = 48000
    pair = ScenePair(oracle, odb, ctx)
    pcms = [synth_pcm(rng, 70000, rate) for _ in range(6)]
    ctls = []
    for i in range(40):
        _, c = pair.play_buffered(rate, pcms[i % 6], 0.0, rand_in_shell(rng, 1, 80), rng.uniform(-20, 20, 3).astype(F32), 0.1,
                                  max_distance=120.0, ring_rate=rate, buffer_duration=0.15,
                                  gain=float(rng.uniform(0.2, 1.0)) if i % 2 else None,
                                  speed=float(rng.uniform(0.8, 1.25)) if i % 3 == 0 else None)
        ctls.append(c)
    for i in range(30):
        pair.play(rate, pcms[i % 6], 1.0, rand_in_shell(rng, 2, 80), rng.uniform(-20, 20, 3).astype(F32))
    for step in range(10):
        if step == 3:
            for c in ctls:
                if "gain" in c:
                    c["gain"][0].control_set_amplitude_ratio(0.3)
                    c["gain"][1].set_amplitude_ratio(0.3)
        if step == 5:
            q = rng.normal(size=4)
            pair.set_listener_rotation((q / np.linalg.norm(q)).astype(F32))
            for i in range(0, 70, 5):
                pair.set_motion(i, rand_in_shell(rng, 1, 80), rng.uniform(-20, 20, 3).astype(F32), bool(i % 2))
        ref, ref64, out = pair.step(rate, 1024 if step % 2 else 512)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)
    assert pair.dev.len(True) == pair.ref.len(True) == 40
    assert pair.dev.len(False) == pair.ref.len(False) == 30
    cnt = pair.dev.last_job_counters()  # buffered sources ride the staged kernel unless their reads wrap this callback
    assert cnt["staged"] + cnt["general"] + cnt["ring_literal"] == 70 and cnt["staged"] >= 60


def test_buffered_ring_wrap_and_finish(oracle, odb, ctx):
    """Small ring (wraps every few callbacks), ring rate different from the PCM rate, and a source that ends:
    dropped once its tail has propagated (spatial.rs:243-261)."""
    rng = np.random.default_rng(62)
    pair = ScenePair(oracle, odb, ctx)
    short = synth_pcm(rng, 6000, 44100)
    pair.play_buffered(44100, short, 0.0, [3.0, 0.0, 1.0], [0.5, 0.0, 0.0], 0.1, max_distance=10.0, ring_rate=48000,
                       buffer_duration=0.05)
    pair.play_buffered(44100, short, -0.02, [0.0, 2.0, 0.0], [0.0, 0.0, 0.0], 0.1, max_distance=5.0, ring_rate=32000,
                       buffer_duration=0.03)
    lens = []
    for _ in range(30):
        ref, ref64, out = pair.step(48000, 512)
        assert_mix_close(out, ref, ref64)
        cursors_equal(pair)
        assert pair.dev.len(True) == pair.ref.len(True)
        lens.append(pair.ref.len(True))
        for hr, hd in zip(pair.ref_handles, pair.dev_handles):
            assert hr.is_finished() == hd.is_finished()
    assert lens[0] == 2 and lens[-1] == 0


def test_multi_second_ring_reads_far_offsets_literally(oracle, odb, ctx):
    """A delay ring of > 2^17 samples (max_distance 1500 m at 48 kHz: 4.5 s): at such absolute offsets the f32 cursor
    chain drifts by more than the staged kernel's window margin, so those reads take the literal ring kernel - and
    stay bit-exact against Ring::sample (ring.rs:51-79)."""
    rng = np.random.default_rng(61)
    rate = 48000
    pcm = synth_pcm(rng, 400000, rate)
    pair = ScenePair(oracle, odb, ctx)
    pair.dev.set_kernel_variant(0)
    pair.play_buffered(rate, pcm, 0.0, [900.0, 50.0, -120.0], [-40.0, 3.0, 4.0], 0.5, max_distance=1500.0, ring_rate=rate,
                       buffer_duration=0.1)
    kinds = set()
    for n in (1024, 2048, 4096) * 30:   # ~4.5 s: the write cursor passes 2^17 and the ring's end
        ref, _, out = pair.step(rate, n)
        np.testing.assert_array_equal(out, ref)
        cursors_equal(pair)
        cnt = pair.dev.last_job_counters()
        kinds.add("staged" if cnt["staged"] else "literal")
    assert np.abs(out).max() > 0
    assert kinds == {"staged", "literal"}
