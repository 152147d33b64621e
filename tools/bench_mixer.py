"""Secondary configs of BASELINE.json on the device Mixer path (not the headline bench):
  c4: 262144 static stereo FramesSignal sources under Gain, Tanh over the mixer, 1024 frames @96 kHz
  c5: 4096 mono Speed<FramesSignal> sources, ratio U[0.5, 2.0), 4096 frames @48 kHz
Prints one JSON line: source-frames/s, ms per callback, algorithmic GB/s and fraction of the measured HBM peak."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oddio_b200 as odb
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4", choices=["c4", "c5"])
ap.add_argument("--sources", type=int, default=0)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = odb.Context(0, stream=stream.cuda_stream)
K, W = a.steps, a.warmup
rng = np.random.default_rng(4 if a.config == "c4" else 5)
if a.config == "c4":
    N, M, rate, ch = a.sources or 262144, 1024, 96000, 2
    L = (K + W + 2) * M + 2048
    speeds = None
else:
    N, M, rate, ch = a.sources or 4096, 4096, 48000, 1
    speeds = rng.uniform(0.5, 2.0, N).astype(np.float32)
    L = int(2.0 * (K + W + 2) * M) + 2048
gen = torch.Generator(device=dev); gen.manual_seed(7)
frames = []
t0 = time.time()
B = 256
kk = torch.arange(L, device=dev, dtype=torch.float32)
for b0 in range(0, N, B):
    nb = min(B, N - b0)
    w = torch.tensor(rng.uniform(100, 4000, (nb, 1, ch)) * 2 * np.pi / rate, device=dev, dtype=torch.float32)
    x = (0.5 * torch.sin(w * kk[None, :, None]) + 0.05 * (2 * torch.rand((nb, L, ch), device=dev, generator=gen) - 1)).contiguous()
    torch.cuda.synchronize(dev)
    for r in range(nb):
        frames.append(odb.Frames.from_device(rate, ch, x[r].data_ptr(), L, ctx))
    del x
with torch.cuda.stream(stream):
    ctl, mixer = odb.Mixer.new(ch, ctx)
    top = mixer
    if a.config == "c4":
        top = odb.Tanh(mixer)
        for i in range(N):
            g = odb.Gain(odb.FramesSignal(frames[i], 0.0)); g.set_amplitude_ratio(float(rng.uniform(0.05, 1.0)) * 1e-3)
            ctl.play(g)
    else:
        for i in range(N):
            sc, sp = odb.Speed.new(odb.FramesSignal(frames[i], 0.0)); sc.set_speed(float(speeds[i]))
            ctl.play(sp)
    setup = time.time() - t0
    tile = torch.zeros((M, ch), device=dev, dtype=torch.float32)
    interval = float(np.float32(1.0) / np.float32(rate))
    for _ in range(W):
        top.sample_device(interval, tile.data_ptr(), M)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        top.sample_device(interval, tile.data_ptr(), M)
    e1.record(stream)
    torch.cuda.synchronize(dev)
ms = e0.elapsed_time(e1) / K
cnt = mixer.last_job_counters()
assert len(mixer) == N, "a source finished during timing"
ds_mean = 1.0 if speeds is None else float(speeds.mean())
alg = 4.0 * ch * M * ds_mean * N
peak, _ = bench.peaks()
print(json.dumps({"config": a.config, "sources": N, "frames": M, "channels": ch, "value": N * M / (ms * 1e-3), "unit": "source-frames/s",
                  "ms_per_step": ms, "alg_GBs_whole_callback": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak_whole_callback": alg / (ms * 1e-3) / 1e9 / peak,
                  "jobs": cnt, "setup_s": round(setup, 1), "checksum": float(tile.abs().sum().item())}))
