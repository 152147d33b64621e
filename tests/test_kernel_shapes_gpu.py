"""The alternative shapes of the one-launch callback kernel (odb_scene_mix.cu, ODB_SMX_CFG in the environment) render
the same scene as the shipped shape.

ODB_SMX_CFG is read once per process, so every shape renders in its own interpreter (with a timeout: a hang of the
producer / consumer hand-over of the warp-specialised shapes must fail the test, not stall the suite). The scene gives
every warp team several batches (the slot ring of the warp-specialised shapes wraps), mixes doppler, static (ds ~= 1
path, frames.rs:180-187) and literal-path sources (FixedGain), and renders a short, a full and a two-tile callback.

  shape 4 (16 consumer + 4 producer warps, setmaxnreg) keeps the shipped kernel's source -> warp assignment and
          accumulation order: bit-identical output;
  shapes 1, 3, 5 sum in another order: within the mix tolerance of the shipped shape.
The shipped shape itself is held against the oracle by the other GPU tests."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RENDER = r"""
import sys
import numpy as np
import oddio_b200 as odb

out_path, n_sources = sys.argv[1], int(sys.argv[2])
rng = np.random.default_rng(11)
rate = 48000
ctx = odb.init(0)
blocks = []
for b in range(16):
    k = np.arange(40000, dtype=np.float64)
    x = 0.5 * np.sin(2 * np.pi * rng.uniform(100.0, 4000.0) * k / rate + rng.uniform(0, 6.28)) + 0.05 * rng.uniform(-1, 1, k.size)
    blocks.append(odb.Frames.from_slice(rate, x.astype(np.float32), ctx))
ctl, scene = odb.SpatialScene.new(ctx)
for i in range(n_sources):
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    pos = (d * rng.uniform(2.0, 100.0)).astype(np.float32)
    kind = i % 16
    vel = np.zeros(3, np.float32) if kind == 3 else rng.uniform(-30, 30, 3).astype(np.float32)
    _, sig = odb.FramesSignal.new(blocks[i % len(blocks)], float(rng.uniform(0.0, 0.3)))
    if kind == 7:
        sig = odb.FixedGain(sig, -6.0)  # the literal path (tail of the kernel)
    ctl.play(sig, odb.SpatialOptions(pos, vel, 0.1))
outs = []
for n in (700, 1024, 2048):
    out = np.zeros((n, 2), dtype=np.float32)
    odb.run(scene, rate, out)
    outs.append(out)
np.save(out_path, np.concatenate(outs))
print(scene.last_job_counters())
"""


def render(tmp_path, cfg, n_sources):
    out = tmp_path / f"shape{cfg}.npy"
    env = dict(os.environ, ODB_SMX_CFG=str(cfg), PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    try:
        r = subprocess.run([sys.executable, "-c", RENDER, str(out), str(n_sources)], env=env, cwd=ROOT, capture_output=True,
                           text=True, timeout=120)
    except subprocess.TimeoutExpired:
        pytest.fail(f"kernel shape {cfg} did not finish (hand-over hang?)")
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


@pytest.fixture(scope="module")
def shipped(tmp_path_factory):
    return render(tmp_path_factory.mktemp("shapes"), 0, 12000)


def test_warp_specialised_shape_is_bit_identical(tmp_path, shipped):
    np.testing.assert_array_equal(render(tmp_path, 4, 12000), shipped)


@pytest.mark.parametrize("cfg", [1, 3, 5])
def test_other_shapes_within_the_mix_tolerance(tmp_path, shipped, cfg):
    out = render(tmp_path, cfg, 12000).astype(np.float64)
    ref = shipped.astype(np.float64)
    rms = float(np.sqrt(np.mean(ref ** 2)))
    tol = 1e-5 * np.maximum(np.abs(ref), rms)
    worst = float(np.max(np.abs(out - ref) / tol))
    assert worst <= 1.0, f"shape {cfg}: worst sample {worst:.3g} x tolerance"
