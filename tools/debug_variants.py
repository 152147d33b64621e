"""Debug aid: one moving source, variant 0 (staged strict) vs variant 1 (general); lists mismatching frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oddio_b200 as odb
from helpers import synth_pcm

ctx = odb.init(0)
rng = np.random.default_rng(1)
pcm = synth_pcm(rng, 60000, 48000)
outs = {}
for variant in (0, 1):
    ctl, scene = odb.SpatialScene.new(ctx)
    scene.set_kernel_variant(variant)
    fr = odb.Frames.from_slice(48000, pcm, ctx)
    ctl.play(odb.FramesSignal(fr, 0.5), odb.SpatialOptions([3.0, 1.0, -2.0], [10.0, -3.0, 4.0], 0.1))
    out = np.zeros((256, 2), np.float32)
    odb.run(scene, 48000, out)
    outs[variant] = out.copy()
    print(variant, scene.last_job_counters())
a, b = outs[0], outs[1]
bad = np.argwhere(a != b)
print(len(bad), "mismatches")
for i, e in bad[:40]:
    print(i, e, a[i, e].view(np.uint32) - b[i, e].view(np.uint32) if False else (a[i, e], b[i, e], int(a[i, e].view(np.int32)) - int(b[i, e].view(np.int32))))
print("frames mod 4 histogram:", np.bincount(bad[:, 0] % 4, minlength=4), "ears:", np.bincount(bad[:, 1], minlength=2))
print("frames:", sorted(set(bad[:, 0].tolist()))[:80])
