"""BASELINE.json's configurations at their FULL sizes on the device, checked through size-independent
properties (the oracle needs minutes for these sizes, so it checks samples of them):

* C3 (SpatialScene, 65 536 moving sources x 1024 frames): source-set linearity (the mix of all sources equals
  the sum of the mixes of two disjoint halves), strict vs FMA variant, the staged kernel takes every job,
  and - for a random sample of the sources - the device's f64 cursors are bit-identical to the CPU oracle's
  and the sample's mix matches the oracle's;
* C5 (Mixer, 4096 Speed sources x 4096 frames): small enough for the oracle - compared directly at full size;
* C4 (Mixer, 262 144 static stereo sources + Gain, Tanh, 96 kHz): linearity over halves, Tanh of the sum, and a
  closed form: every source is on FramesSignal's ds == 1 path (frames.rs:180-187), so the mix is
  sum_i gain_i * lerp(pcm_i[base_i + k], pcm_i[base_i + k + 1], fract_i), evaluated in numpy f64.
"""
import numpy as np
import pytest

from helpers import F32, MixerPair, rand_in_shell, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odb():
    import oddio_b200

    return oddio_b200


@pytest.fixture(scope="module")
def ctx(odb):
    return odb.init(0)


def close_sums(a, b, rel=1e-5):
    """Two f32 sums of the same terms in different orders: per sample within rel * max(|x|, RMS)."""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    rms = float(np.sqrt(np.mean(b64 ** 2)))
    tol = rel * np.maximum(np.abs(b64), rms) + 1e-30
    bad = np.abs(a64 - b64) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} samples off; worst {np.max(np.abs(a64 - b64) / tol):.3g}x tolerance"


def test_c3_full_size(oracle, odb, ctx):
    N, M, RATE, K = 65536, 1024, 48000, 2
    rng = np.random.default_rng(303)
    pcms = [synth_pcm(rng, RATE + int(1.2 * M * (K + 1)) + 2048, RATE) for _ in range(64)]
    frames = [odb.Frames.from_slice(RATE, p, ctx) for p in pcms]
    u = rng.normal(size=(N, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    pos = (u * rng.uniform(2.0, 300.0, (N, 1))).astype(F32)
    vel = rng.uniform(-30, 30, (N, 3)).astype(F32)

    def scene_of(ids, variant):
        ctl, sc = odb.SpatialScene.new(ctx)
        sc.set_kernel_variant(variant)
        controls = []
        for i in ids:
            c, sig = odb.FramesSignal.new(frames[i % 64], 1.0)
            ctl.play(sig, odb.SpatialOptions(pos[i], vel[i], 0.1))
            controls.append(c)
        return ctl, sc, controls

    ids_all = np.arange(N)
    sample = np.sort(rng.choice(N, 48, replace=False))
    scenes = {"all": scene_of(ids_all, 2), "even": scene_of(ids_all[::2], 2), "odd": scene_of(ids_all[1::2], 2),
              "strict": scene_of(ids_all, 0), "sample": scene_of(sample, 0)}
    ref = oracle.SpatialScene()
    ofr = [oracle.Frames.from_slice(RATE, p) for p in pcms]
    osig = []
    for i in sample:
        s = oracle.FramesSignal(ofr[i % 64], 1.0)
        ref.play(s, pos[i], vel[i], 0.1)
        osig.append(s)
    for _ in range(K):
        out = {}
        for name, (_, sc, _) in scenes.items():
            o = np.zeros((M, 2), F32)
            odb.run(sc, RATE, o)
            out[name] = o
        assert float(np.abs(out["all"]).max()) > 1e-2   # a non-trivial mix
        close_sums(out["all"], out["even"].astype(np.float64) + out["odd"].astype(np.float64))  # linearity over the source set
        close_sums(out["all"], out["strict"])                                                   # FMA-contracted vs strict values
        for name in ("all", "strict"):
            assert scenes[name][1].last_job_counters() == {"general": 0, "staged": N, "resampled": 0, "ring_literal": 0}
        r = oracle.run(ref, RATE, M)
        close_sums(out["sample"], r)
        # f64 cursors of the sampled sources: bit-identical to the oracle's, in the full-size scenes too
        for j, i in enumerate(sample):
            t_ref = osig[j].t
            assert scenes["sample"][2][j].cursor()[0] == t_ref
            assert scenes["all"][2][i].cursor()[0] == t_ref
            assert scenes["strict"][2][i].cursor()[0] == t_ref
    for _, sc, _ in scenes.values():
        assert sc.len() in (N, N // 2, len(sample))
        sc.close()


def test_c5_full_size_against_the_oracle(oracle, odb, ctx):
    rng = np.random.default_rng(505)
    rate, N, M = 48000, 4096, 4096
    pair = MixerPair(oracle, odb, ctx, 1)
    pcms = [synth_pcm(rng, 2 * M * 3 + 4096, rate, 1) for _ in range(16)]
    for i in range(N):
        pair.play(rate, pcms[i % 16], 0.0, speed=float(rng.uniform(0.5, 2.0)))
    for _ in range(2):
        ref, ref64, out = pair.step(rate, M)
        close_sums(out, ref)
        close_sums(out, ref64)
    cnt = pair.dev_mixer.last_job_counters()
    assert cnt["general"] == 0 and cnt["resampled"] == 4 * N
    for it in pair.items[:: N // 64]:
        assert it["dev_frames_control"].cursor()[0] == it["ref_frames_signal"].t


def test_c4_full_size(odb, ctx):
    rng = np.random.default_rng(404)
    rate, N, M, K = 96000, 262144, 1024, 2
    L = M * (K + 1) + 64
    pcms = [synth_pcm(rng, L, rate, 2) for _ in range(16)]
    frames = [odb.Frames.from_slice(rate, p, ctx) for p in pcms]
    gains = (rng.uniform(0.05, 1.0, N) * 2e-3).astype(F32)
    starts = rng.integers(0, 32, N)  # whole frames: fract == 0, base = start + k

    def mixer_of(ids, tanh):
        ctl, mx = odb.Mixer.new(2, ctx)
        sig = odb.Tanh(mx) if tanh else mx
        for i in ids:
            _, s = odb.FramesSignal.new(frames[i % 16], float(starts[i]) / rate)
            _, g = odb.Gain.new(s)
            g.set_amplitude_ratio(float(gains[i]))
            ctl.play(g)
        return sig, mx

    ids = np.arange(N)
    full, full_mx = mixer_of(ids, False)
    tanh, _ = mixer_of(ids[::2], True)
    even, _ = mixer_of(ids[::2], False)
    odd, _ = mixer_of(ids[1::2], False)
    # closed form in f64: per PCM block, the gains of its sources binned by start frame
    w = np.zeros((16, 32))
    np.add.at(w, (ids % 16, starts), gains.astype(np.float64))
    for k in range(K):
        outs = []
        for sig in (full, tanh, even, odd):
            o = np.zeros((M, 2), F32)
            odb.run(sig, rate, o)
            outs.append(o)
        o_full, o_tanh, o_even, o_odd = outs
        want = np.zeros((M, 2))
        for b in range(16):
            for s0 in range(32):
                if w[b, s0] != 0.0:
                    want += w[b, s0] * pcms[b][s0 + k * M: s0 + (k + 1) * M].astype(np.float64)
        assert float(np.abs(want).max()) > 1e-2
        close_sums(o_full, want)
        close_sums(o_full, o_even.astype(np.float64) + o_odd.astype(np.float64))
        np.testing.assert_allclose(o_tanh, np.tanh(o_even.astype(np.float64)), rtol=0, atol=4e-7)  # tanh.rs:26 on the sum
        assert full_mx.last_job_counters() == {"general": 0, "staged": N, "resampled": 0, "ring_literal": 0}
