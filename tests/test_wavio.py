"""WAV ingest / offline render either side of the hot path (examples/wav.rs, examples/offline.rs)."""
import os
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quantize_matches_rust_as_cast():
    from oddio_b200.wavio import quantize_i16

    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -2.0, 0.99999, np.nan, 3.0518e-5, -3.0518e-5], dtype=np.float32)
    q = quantize_i16(x)
    # (s * 32767.0) as i16: truncation toward zero, saturation, NaN -> 0
    assert q.tolist() == [0, 16383, -16383, 32767, -32767, 32767, -32768, 32766, 0, 0, 0]


def test_wav_roundtrip_and_scaling(tmp_path):
    from oddio_b200.wavio import quantize_i16, read_wav, render_offline

    rng = np.random.default_rng(0)
    blocks = [rng.uniform(-1, 1, (64, 2)).astype(np.float32) for _ in range(5)]
    it = iter(blocks)
    path = str(tmp_path / "t.wav")
    assert render_offline(lambda m: next(it), path, 22050, 64, 5) == 320
    rate, x = read_wav(path)
    assert rate == 22050 and x.shape == (320, 2)
    want = np.concatenate([quantize_i16(b) for b in blocks]).astype(np.float32) / np.float32(32767.0)
    np.testing.assert_array_equal(x, want)  # examples/wav.rs:33: sample as f32 / (2^15 - 1) as f32


def test_reads_the_references_example_file():
    from oddio_b200.wavio import read_wav

    p = "/root/reference/examples/wav/stereo-test.wav"
    if not os.path.exists(p):
        pytest.skip("reference tree not present on this box")
    rate, x = read_wav(p)
    assert x.ndim == 2 and x.shape[1] == 2 and rate > 0 and np.abs(x).max() <= 1.0 + 1e-6


@pytest.mark.gpu
def test_offline_example_matches_oracle_render(tmp_path, oracle):
    """examples/offline.rs rendered by the device path and by the oracle: one source, so the f32 blocks - and
    therefore the 16-bit files - are identical."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import offline

    dev_path, ref_path, fma_path = str(tmp_path / "dev.wav"), str(tmp_path / "ref.wav"), str(tmp_path / "fma.wav")
    offline.main(dev_path, kernel_variant=0)   # strict arithmetic: bit-identical f32 blocks
    offline.main(fma_path)                     # the library default (FMA-contracted values): within one 16-bit step
    from oddio_b200 import wavio

    scene = oracle.SpatialScene()
    frames = oracle.Frames.from_slice(offline.RATE, offline.boop())
    scene.play(oracle.FramesSignal(frames, 0.0), [-offline.SPEED, 10.0, 0.0], [offline.SPEED, 0.0, 0.0], 0.1)
    wavio.render_offline(lambda m: oracle.run(scene, offline.RATE, m), ref_path, offline.RATE, offline.BLOCK_SIZE,
                         offline.RATE * offline.DURATION_SECS // offline.BLOCK_SIZE)
    with wave.open(dev_path) as a, wave.open(ref_path) as b:
        assert a.getnframes() == b.getnframes() == 258 * 512
        assert a.readframes(a.getnframes()) == b.readframes(b.getnframes())
    (_, xf), (_, xr) = wavio.read_wav(fma_path), wavio.read_wav(ref_path)
    assert xf.shape == xr.shape and float(np.abs(xf - xr).max()) <= 1.01 / 32767.0


@pytest.mark.gpu
def test_device_int16_ingest_is_bit_identical_to_the_host_scaling(tmp_path):
    """examples/wav.rs:30-37 on the device (odb_frames_from_i16) against read_wav's numpy statement of it: the same
    f32 values, read back through a unit-rate Mixer (every frame on FramesSignal's ds == 1 path with fract 0)."""
    import oddio_b200 as odb
    from oddio_b200 import wavio

    ctx = odb.init(0)
    rng = np.random.default_rng(5)
    for ch, width in ((2, 2), (1, 2), (2, 1)):
        n, rate = 1000, 22050  # one mixer chunk: t = 0, so base = k and fract = 0 exactly
        if width == 2:
            ints = rng.integers(-32768, 32768, (n, ch)).astype("<i2")
            ints[:4] = np.array([[-32768] * ch, [32767] * ch, [0] * ch, [-1] * ch], dtype="<i2")
            raw = ints.tobytes()
        else:
            raw = rng.integers(0, 256, (n, ch)).astype(np.uint8).tobytes()
        path = str(tmp_path / f"in_{ch}_{width}.wav")
        with wave.open(path, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(width); w.setframerate(rate); w.writeframes(raw)
        r, want = wavio.read_wav(path)
        fr = wavio.frames_from_wav(path, ctx)          # int16 upload + device scaling
        ctl, mixer = odb.Mixer.new(ch, ctx)
        ctl.play(odb.FramesSignal(fr, 0.0))
        got = odb.run(mixer, rate, np.zeros((n, ch) if ch > 1 else (n,), np.float32))
        np.testing.assert_array_equal(got[: n - 1], want.reshape(got.shape)[: n - 1])  # (the last frame lerps against the end)
        mixer.close()


@pytest.mark.gpu
def test_device_quantisation_matches_the_rust_cast(tmp_path, oracle):
    """odb_*_sample_i16 against quantize_i16 (the numpy statement of `(sample * i16::MAX as f32) as i16`) on the f32
    blocks of an identical twin scene; includes saturation (a loud source) and the Tanh epilogue."""
    import oddio_b200 as odb
    from oddio_b200 import wavio

    ctx = odb.init(0)
    rate, n = 44100, 512
    t = np.arange(rate, dtype=np.float32) / np.float32(rate)
    pcm = (np.sin(t * np.float32(500.0 * 2.0 * np.pi)) * np.float32(80.0)).astype(np.float32)  # examples/offline.rs:12
    fr = odb.Frames.from_slice(rate, pcm, ctx)
    for wrap in (None, odb.Tanh):
        twins = []
        for _ in range(2):
            ctl, scene = odb.SpatialScene.new(ctx)
            ctl.play(odb.FramesSignal(fr, 0.0), odb.SpatialOptions([-50.0, 10.0, 0.0], [50.0, 0.0, 0.0], 0.1))
            ctl.play(odb.FramesSignal(fr, 0.1), odb.SpatialOptions([0.5, 0.2, 0.0], [0.0, 0.0, 0.0], 0.1))  # close: clips
            twins.append(wrap(scene) if wrap else scene)
        interval = float(np.float32(1.0) / np.float32(rate))
        saturated = 0
        for _ in range(6):
            f32 = twins[0].sample(interval, n)
            i16 = twins[1].sample_i16(interval, n)
            want = wavio.quantize_i16(f32)
            np.testing.assert_array_equal(i16, want)
            saturated += int((np.abs(want.astype(np.int32)) >= 32767).sum())
        assert wrap is not None or saturated > 0
    # and through the file writer
    ctl, scene = odb.SpatialScene.new(ctx)
    ctl.play(odb.FramesSignal(fr, 0.0), odb.SpatialOptions([-50.0, 10.0, 0.0], [50.0, 0.0, 0.0], 0.1))
    assert wavio.render_offline_device(scene, str(tmp_path / "o.wav"), rate, n, 4) == 4 * n
    r, x = wavio.read_wav(str(tmp_path / "o.wav"))
    assert r == rate and x.shape == (4 * n, 2)
