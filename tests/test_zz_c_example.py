"""The drop-in boundary from compiled code: examples/offline.c (examples/offline.rs through the C ABI, plain C11)
must compile against include/oddio_b200.h with warnings as errors and link against the in-tree library. Without a
GPU it has to fail loudly (no CPU fallback); on the GPU box its WAV file is compared with the same render through
the Python mirror."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def offline_binary(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from oddio_b200 import build

    build.build()
    exe = str(tmp_path_factory.mktemp("cex") / "offline_c")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "offline.c"), "-L", os.path.join(ROOT, "oddio_b200"), "-loddio_b200", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "oddio_b200") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe, *args], capture_output=True, text=True, env=env, timeout=300)


def test_c_example_builds_and_has_no_cpu_fallback(offline_binary, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = _run(offline_binary, str(tmp_path / "o.wav"))
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_renders_like_the_python_mirror(offline_binary, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import offline

    import oddio_b200 as odb
    from oddio_b200 import wavio

    c_path, py_path = str(tmp_path / "c.wav"), str(tmp_path / "py.wav")
    r = _run(offline_binary, c_path)
    assert r.returncode == 0, r.stderr
    ctx = odb.init(0)
    frames = odb.Frames.from_slice(offline.RATE, offline.boop(), ctx)
    ctl, scene = odb.SpatialScene.new(ctx)
    ctl.play(odb.FramesSignal(frames, 0.0), odb.SpatialOptions([-offline.SPEED, 10.0, 0.0], [offline.SPEED, 0.0, 0.0], 0.1))
    n_blocks = offline.RATE * offline.DURATION_SECS // offline.BLOCK_SIZE
    wavio.render_offline_device(scene, py_path, offline.RATE, offline.BLOCK_SIZE, n_blocks)
    (rc, xc), (rp, xp) = wavio.read_wav(c_path), wavio.read_wav(py_path)
    assert rc == rp == offline.RATE and xc.shape == xp.shape == (n_blocks * offline.BLOCK_SIZE, 2)
    assert float(np.abs(xp).max()) > 0.01
    # the two programs synthesise the boop with different sine routines (glibc sinf / numpy): within two 16-bit steps
    assert float(np.abs(xc - xp).max()) <= 2.01 / 32767.0
