"""C3b: the C3 sources through play_buffered(max_distance=350, rate=48000, buffer_duration=0.1).
Prints source-frames/s and ms per callback after the rings have filled (warm-up >= max_delay)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oddio_b200 as odb
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--sources", type=int, default=16384)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=56)
a = ap.parse_args()
N, M, K, W = a.sources, 1024, a.steps, a.warmup
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = odb.Context(0, stream=stream.cuda_stream)
pos, vel, freq, phase = bench.scene_geometry(N)
L = int(1.0 * M * (K + W + 2)) + 4096
gen = torch.Generator(device=dev); gen.manual_seed(3)
kk = torch.arange(L, device=dev, dtype=torch.float32)
frames = []
for b0 in range(0, N, 512):
    ids = np.arange(b0, min(N, b0 + 512))
    w = torch.tensor(2 * np.pi * freq[ids] / 48000, device=dev, dtype=torch.float32)[:, None]
    x = (0.5 * torch.sin(w * kk[None, :]) + 0.05 * (2 * torch.rand((len(ids), L), device=dev, generator=gen) - 1)).contiguous()
    torch.cuda.synchronize(dev)
    for r in range(len(ids)):
        frames.append(odb.Frames.from_device(48000, 1, x[r].data_ptr(), L, ctx))
    del x
with torch.cuda.stream(stream):
    ctl, scene = odb.SpatialScene.new(ctx)
    for i in range(N):
        ctl.play_buffered(odb.FramesSignal(frames[i], 0.0), odb.SpatialOptions(pos[i], vel[i], 0.1), 350.0, 48000, 0.1)
    tile = torch.zeros((M, 2), device=dev, dtype=torch.float32)
    interval = float(np.float32(1.0) / np.float32(48000))
    for _ in range(W):
        scene.sample_device(interval, tile.data_ptr(), M)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        scene.sample_device(interval, tile.data_ptr(), M)
    e1.record(stream)
    torch.cuda.synchronize(dev)
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": "c3b buffered", "sources": N, "frames": M, "value": N * M / (ms * 1e-3), "ms_per_step": ms,
                  "len": scene.len(True), "checksum": float(tile.abs().sum().item())}))
