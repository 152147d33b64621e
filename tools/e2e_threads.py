"""End-to-end callback time with the control thread of bench.py, varying one thing at a time (developer tool).
usage: e2e_threads.py [inline|thread] [own|torch]"""
import os, sys, time, threading, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oddio_b200 as odb
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "thread"
strm = sys.argv[2] if len(sys.argv) > 2 else "torch"
N, M, K, W = 65536, 1024, 16, 3
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = odb.Context(0, stream=stream.cuda_stream) if strm == "torch" else odb.Context(0)
pos, vel, freq, phase = bench.scene_geometry(N)
L = bench.pcm_len(M, 40)
x = (0.5 * torch.sin(torch.arange(L, device=dev, dtype=torch.float32) * 0.05)).contiguous()
torch.cuda.synchronize()
fr = odb.Frames.from_device(48000, 1, x.data_ptr(), L, ctx)
ctl, scene = odb.SpatialScene.new(ctx)
scene.set_kernel_variant(2)
hs = [ctl.play(odb.FramesSignal(fr, 1.0), odb.SpatialOptions(pos[i], vel[i], 0.1)) for i in range(N)]
out = np.zeros((M, 2), np.float32)
n_upd = N // 16
rng = np.random.default_rng(7)
upd = []
for s in range(W + K):
    sel = rng.choice(N, n_upd, replace=False)
    ids = (C.c_uint64 * n_upd)(*[hs[i]._src for i in sel])
    upd.append((ids, pos[sel].copy(), vel[sel].copy()))
go = threading.Semaphore(0)

def control_thread():
    for s in range(W + K):
        go.acquire()
        ids, p, v = upd[s]
        ctl.set_motion_ids(ids, n_upd, p, v)

def step(s):
    if mode == "thread":
        go.release()
    else:
        ids, p, v = upd[s]
        ctl.set_motion_ids(ids, n_upd, p, v)
    odb.run(scene, 48000, out)

th = threading.Thread(target=control_thread, daemon=True)
if mode == "thread":
    th.start()
for s in range(W):
    step(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(W, W + K):
    step(s)
if mode == "thread":
    th.join()
torch.cuda.synchronize()
print(mode, strm, "us per callback:", round((time.perf_counter() - t0) / K * 1e6, 1))
scene.close()
