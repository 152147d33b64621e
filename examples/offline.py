"""examples/offline.rs on the device path: a 500 Hz boop flying past the listener at 50 m/s, rendered to offline.wav.

    python examples/offline.py [out.wav]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oddio_b200 as odb
from oddio_b200 import wavio

DURATION_SECS, RATE, BLOCK_SIZE, SPEED = 3, 44100, 512, 50.0


def boop() -> np.ndarray:
    i = np.arange(RATE * DURATION_SECS, dtype=np.float32)
    t = i / np.float32(RATE)
    return (np.sin(t * np.float32(500.0) * np.float32(2.0) * np.float32(np.pi), dtype=np.float32) * np.float32(80.0)).astype(np.float32)


def main(path: str = "offline.wav", kernel_variant=None) -> None:
    """`kernel_variant=0` selects the strict arithmetic (every value operation unfused, as Rust computes it): one source,
    so the file is then bit-identical to the reference's; the default (FMA-contracted values) is within one 16-bit step."""
    ctx = odb.init(0)
    frames = odb.Frames.from_slice(RATE, boop(), ctx)
    scene_handle, scene = odb.SpatialScene.new(ctx)
    if kernel_variant is not None:
        scene.set_kernel_variant(kernel_variant)
    scene_handle.play(odb.FramesSignal(frames, 0.0), odb.SpatialOptions([-SPEED, 10.0, 0.0], [SPEED, 0.0, 0.0], 0.1))
    block = np.zeros((BLOCK_SIZE, 2), dtype=np.float32)
    n = wavio.render_offline(lambda m: odb.run(scene, RATE, block), path, RATE, BLOCK_SIZE, RATE * DURATION_SECS // BLOCK_SIZE)
    print(f"wrote {n} frames to {path}")


if __name__ == "__main__":
    main(*sys.argv[1:2])
