"""BASELINE.json's configurations at their FULL sizes on the device.

* `test_c3_full_size_against_the_oracle`, `test_c4_full_size_against_the_oracle`: the whole scene / mixer in the CPU
  oracle too (one thread: its reference-order f32 sum is the reference's), every source with its OWN PCM block,
  compared with SURVEY.md section 7 H4 (iii): device and reference-order f32 both against the f64-accumulated truth;
* the older tests below check the same sizes through size-independent properties:

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


def h4_large(dev, ref32, ref64, rel=1e-5):
    """SURVEY.md section 7 H4 (iii), for source counts where the reference's own sequential f32 sum is > 1e-5 away
    from the truth: the device's error against the f64-accumulated oracle is within rel * max(|truth|, RMS) per sample
    and no larger (RMS over the buffer) than the reference-order f32 sum's own error. Returns the three figures."""
    dev, ref32, ref64 = (np.asarray(x, np.float64) for x in (dev, ref32, ref64))
    rms = float(np.sqrt(np.mean(ref64 ** 2)))
    assert rms > 1e-3, "trivial mix"
    tol = rel * np.maximum(np.abs(ref64), rms)
    e_dev, e_ref = np.abs(dev - ref64), np.abs(ref32 - ref64)
    bad = e_dev > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} samples off against the f64 truth; worst {np.max(e_dev / tol):.3g}x tolerance"
    rms_dev, rms_ref = float(np.sqrt(np.mean(e_dev ** 2))), float(np.sqrt(np.mean(e_ref ** 2)))
    assert rms_dev <= 1.05 * rms_ref + 1e-12, f"device error {rms_dev:.3g} (RMS) exceeds the reference-order sum's own {rms_ref:.3g}"
    return {"dev_vs_truth_max_rel_rms": float(e_dev.max() / rms), "ref32_vs_truth_max_rel_rms": float(e_ref.max() / rms),
            "dev_vs_ref32_max_rel_rms": float(np.abs(dev - ref32).max() / rms)}


def test_c3_full_size_against_the_oracle(oracle, odb, ctx):
    """C3 as bench.py runs it - 65 536 moving sources, every one with its own PCM block, 1024-frame callbacks - in
    the oracle as one scene on one thread (the reference's order of additions) and on the device."""
    import torch

    N, M, RATE, K = 65536, 1024, 48000, 2
    rng = np.random.default_rng(3030)
    u = rng.normal(size=(N, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    dist = rng.uniform(2.0, 300.0, (N, 1))
    pos = (u * dist).astype(F32)
    v = rng.normal(size=(N, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    vel = (v * rng.uniform(0.0, 50.0, (N, 1))).astype(F32)
    # Every source plays a block of its own that holds just what the K callbacks can touch: block i starts `off[i]`
    # frames into the sound, and its FramesSignal starts at 1.0 s - off[i] / rate (frames.rs:156), so the first read
    # (1.0 s minus the propagation delay) lands ~128 frames into the block. Same (block, start) on both sides.
    off = np.floor(RATE * (1.0 - dist[:, 0] / 343.0)).astype(np.int64) - 128
    start = 1.0 - off.astype(np.float64) / RATE
    L = 128 + int(1.2 * M * K) + 512
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(33)
    kk = torch.arange(L, device=dev, dtype=torch.float32)
    w = torch.tensor(2 * np.pi * rng.uniform(100.0, 4000.0, N) / RATE, device=dev, dtype=torch.float32)[:, None]
    ph = torch.tensor(rng.uniform(0, 2 * np.pi, N), device=dev, dtype=torch.float32)[:, None]
    ctl, sc = odb.SpatialScene.new(ctx)
    ref = oracle.SpatialScene()
    controls, osig, keep = [], [], []
    B = 8192
    for b0 in range(0, N, B):
        x = (0.5 * torch.sin(w[b0:b0 + B] * kk[None, :] + ph[b0:b0 + B])
             + 0.05 * (2 * torch.rand((B, L), device=dev, generator=gen) - 1)).contiguous()
        torch.cuda.synchronize()
        xh = x.cpu().numpy()
        for r in range(B):
            i = b0 + r
            fd = odb.Frames.from_device(RATE, 1, x[r].data_ptr(), L, ctx)
            c, sig = odb.FramesSignal.new(fd, float(start[i]))
            ctl.play(sig, odb.SpatialOptions(pos[i], vel[i], 0.1))
            controls.append(c)
            fo = oracle.Frames.from_slice(RATE, xh[r])
            so = oracle.FramesSignal(fo, float(start[i]))
            ref.play(so, pos[i], vel[i], 0.1)
            osig.append(so)
            keep.append(fo)
        del x
    sample = rng.choice(N, 256, replace=False)
    for k in range(K):
        r32 = oracle.run(ref, RATE, M)
        r64 = ref.out64(M)
        out = np.zeros((M, 2), F32)
        odb.run(sc, RATE, out)
        figures = h4_large(out, r32, r64)
        print(f"C3 full size, callback {k}: {figures}")
        assert sc.last_job_counters() == {"general": 0, "staged": N, "resampled": 0, "ring_literal": 0}
        for i in sample:  # frame cursors: bit-identical
            assert controls[i].cursor()[0] == osig[i].t
    assert sc.len() == N
    sc.close()


def test_c4_full_size_against_the_oracle(oracle, odb, ctx):
    """C4 - 262 144 static stereo sources under Gain, Tanh on the sum, 96 kHz - in the oracle and on the device:
    fractional start positions (the constant-fraction lerp of frames.rs:183-187 at scale), own PCM blocks, and a Gain
    transition (gain.rs:118-121, the literal kernel) on 1/64 of the sources before the second callback."""
    rng = np.random.default_rng(4040)
    rate, N, M, K = 96000, 262144, 1024, 2
    L = M * K + 96
    pair = MixerPair(oracle, odb, ctx, 2, epilogue="tanh")
    block = synth_pcm(rng, L + 4096 + 8, rate, 2)
    gains = rng.uniform(0.05, 1.0, N) * 2e-3
    starts = rng.uniform(0.0, 32.0, N) / rate          # fractional frames
    # 4096 distinct PCM blocks (own blocks for all 262 144 sources would be 4.4 GB of host copies through ctypes)
    blocks = [np.ascontiguousarray(block[o:o + L]) for o in range(0, 4096)]
    for i in range(N):
        pair.play(rate, blocks[i % 4096], float(starts[i]), gain=float(gains[i]))
    for k in range(K):
        if k == 1:
            for i in range(0, N, 64):
                pair.set_gain_ratio(i, float(gains[i] * 0.5))
        ref, ref64, out = pair.step(rate, M)
        # the oracle's out64 is the mixer's sum before Tanh; compare the epilogue through the reference's f32 output
        truth = np.tanh(ref64)
        figures = h4_large(out, ref, truth, rel=1e-5)
        print(f"C4 full size, callback {k}: {figures}")
        cnt = pair.dev_mixer.last_job_counters()
        assert cnt["general"] == (0 if k == 0 else N // 64) and cnt["staged"] == (N if k == 0 else N - N // 64)
    for it in pair.items[:: N // 128]:
        assert it["dev_frames_control"].cursor()[0] == it["ref_frames_signal"].t


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
