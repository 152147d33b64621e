"""Times the pieces of one end-to-end callback (host buffers) on C3-like sizes."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oddio_b200 as odb
import bench

N, M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 1024
dev = torch.device("cuda", 0)
ctx = odb.Context(0)
pos, vel, freq, phase = bench.scene_geometry(N)
L = bench.pcm_len(M, 40)
x = (0.5 * torch.sin(torch.arange(L, device=dev, dtype=torch.float32) * 0.05)).contiguous()
torch.cuda.synchronize()
fr = odb.Frames.from_device(48000, 1, x.data_ptr(), L, ctx)   # shared PCM: this script times the host side
ctl, scene = odb.SpatialScene.new(ctx)
scene.set_kernel_variant(2)
hs = [ctl.play(odb.FramesSignal(fr, 1.0), odb.SpatialOptions(pos[i], vel[i], 0.1)) for i in range(N)]
out = np.zeros((M, 2), np.float32)
n_upd = N // 16
ids = (C.c_uint64 * n_upd)(*[hs[i]._src for i in range(0, N, 16)])
p = pos[::16].copy(); v = vel[::16].copy()
for _ in range(3):
    ctl.set_motion_ids(ids, n_upd, p, v); odb.run(scene, 48000, out)
t = {"set_motion": 0.0, "run": 0.0}
K = 20
for _ in range(K):
    a = time.perf_counter(); ctl.set_motion_ids(ids, n_upd, p, v); b = time.perf_counter(); odb.run(scene, 48000, out); c = time.perf_counter()
    t["set_motion"] += b - a; t["run"] += c - b
print({k: round(v / K * 1e6, 1) for k, v in t.items()}, "us per callback;", N, "sources,", n_upd, "updates")
a = time.perf_counter()
for _ in range(K): odb.run(scene, 48000, out)
print("run without updates:", round((time.perf_counter() - a) / K * 1e6, 1), "us")
