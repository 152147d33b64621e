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

    dev_path, ref_path = str(tmp_path / "dev.wav"), str(tmp_path / "ref.wav")
    offline.main(dev_path)
    from oddio_b200 import wavio

    scene = oracle.SpatialScene()
    frames = oracle.Frames.from_slice(offline.RATE, offline.boop())
    scene.play(oracle.FramesSignal(frames, 0.0), [-offline.SPEED, 10.0, 0.0], [offline.SPEED, 0.0, 0.0], 0.1)
    wavio.render_offline(lambda m: oracle.run(scene, offline.RATE, m), ref_path, offline.RATE, offline.BLOCK_SIZE,
                         offline.RATE * offline.DURATION_SECS // offline.BLOCK_SIZE)
    with wave.open(dev_path) as a, wave.open(ref_path) as b:
        assert a.getnframes() == b.getnframes() == 258 * 512
        assert a.readframes(a.getnframes()) == b.readframes(b.getnframes())
