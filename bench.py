#!/usr/bin/env python
"""bench.py — source-frames/s of the spatial-mix hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C3b|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3"): one SpatialScene with 65 536 moving point sources
(`SpatialSceneControl::play` of a FramesSignal each: propagation delay, doppler resampling, distance
attenuation, stereo pan), 1024 stereo frames per callback at 48 kHz. A "step" is one `oddio::run`
callback over all sources. Every source owns its own PCM (no sharing; 19 GB in HBM in total), every
callback reads fresh PCM, so the inputs of a step are larger than L2 by construction.

One JSON line on stdout (rank 0). `value` = N_sources * frames / device time per callback with
everything resident in HBM (CUDA events on the launch stream, max over ranks); `e2e` = the same through the
host-buffer C-ABI call, one callback at a time (`odb_scene_run`: host output tile, D2H inside the timed region, plus
`set_motion` updates for 1/16 of the sources every callback from a second thread, H2D inside the timed region) -
at N = 1 driven by the compiled C harness tools/e2e_native.c (the Python-driven figure is `e2e.python_driven_value`);
`roofline` = algorithmic PCM bytes of the callback kernel / its device time against the measured HBM peak;
`cpu_baseline` = the CPU oracle (the only runnable statement of the Rust reference here) on a bounded sample.

N > 1: one process per GPU. Default `--scaling strong`: the fixed scene of --sources sources is split over the ranks
(round-robin) and EVERY callback ends with the sum of the ranks' tiles, done from inside the callback kernel over
NVLink peer memory (`odb_scene_sample_exchange`, `--lag 1`: callback k receives the sum of callback k - 1, the
pipelined-renderer shape; the `e2e` pass uses lag 0 and writes the summed tile into pinned host memory). The line
then also carries `parity` (a 4096-source sharded scene after the timing: exchanged tile == rank-order sum of the
gathered per-rank tiles bit for bit, identical on all ranks, within tolerance of the CPU oracle) and `extra`
(`weak`: --sources per GPU; `strong_reduce_every_8`: the offline shape with the stand-alone exchange kernels).
`--exchange peer|nccl|none` select the stand-alone kernels, torch.distributed, or no exchange (diagnostic).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RATE = 48000
START_S = 1.0          # FramesSignal::new(frames, 1.0): > 300 m / 343 m/s, so every read hits valid PCM
SHELL = (2.0, 300.0)   # source distance from the listener, metres
SPEED_MAX = 50.0       # m/s  (examples/offline.rs:4 uses 50 m/s)
DS_MAX = 1.0 + SPEED_MAX / 343.0 + 0.01


_REAL_STDOUT = None


def emit(obj):
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def ncu_traffic(kernel: str, n_local: int):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/),
    scaled to this run's source count; None if no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f).get(kernel)
    if not t:
        return None
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) * n_local / 65536.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def scene_geometry(n_sources: int, seed: int = 0x0DD10):
    """Positions uniform in direction, U[2,300] m in range; velocities uniform in direction, U[0,50] m/s."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_sources, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pos = (d * rng.uniform(SHELL[0], SHELL[1], size=(n_sources, 1))).astype(np.float32)
    v = rng.normal(size=(n_sources, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    vel = (v * rng.uniform(0.0, SPEED_MAX, size=(n_sources, 1))).astype(np.float32)
    f = rng.uniform(100.0, 4000.0, size=n_sources)
    ph = rng.uniform(0.0, 2 * np.pi, size=n_sources)
    return pos, vel, f, ph


def pcm_len(frames: int, callbacks: int) -> int:
    return int(START_S * RATE + np.ceil(DS_MAX * frames * (callbacks + 2)) + 2048)


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU while the timed region runs: an `nvidia-smi -lms 100` child
    started before the region (rank 0 only: NVML queries take driver locks, and one poller per rank on an 8-GPU box
    measurably slows the launches of all of them). The timed region of the default run is a few milliseconds, so
    the poller may not land a sample inside it on a box where nvidia-smi starts slowly; `after()` then takes
    readings through NVML right after the region's closing synchronisation (reported as "sampled": "after")."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int, enabled: bool = True):
        self.device, self.rows, self.proc, self.enabled, self.how = device, [], None, enabled, "during"
        self.all_rows, self.t0, self.t1 = [], None, None

    def start(self):
        import atexit

        atexit.register(self.stop)
        return self._start()

    def _start(self):
        """Starts the poller. Called long before the timed region: nvidia-smi needs a noticeable fraction of a second
        to initialise on a multi-GPU box and holds driver locks while it does, which must not overlap the region."""
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._smi_id()}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _smi_id(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",")]
            if self.device < len(ids):
                return ids[self.device]
        return str(self.device)

    def _pump(self):
        for line in self.proc.stdout:
            self.all_rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def __enter__(self):  # the timed region starts
        self.t0 = time.time()
        return self

    def __exit__(self, *a):  # ... and has ended (after its closing synchronisation)
        self.t1 = time.time()
        if self.proc:
            time.sleep(0.12)  # one more polling period: the reading that covers the end of the region
            self.rows = [r for t, r in self.all_rows if self.t0 - 0.02 <= t <= self.t1 + 0.12]
        self.after()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def after(self):
        """No sample landed inside the region: read the clocks now (the device has just finished it)."""
        if not self.enabled or self.rows:
            return
        try:
            import pynvml as p

            p.nvmlInit()
            sid = self._smi_id()
            h = p.nvmlDeviceGetHandleByIndex(int(sid)) if sid.isdigit() else p.nvmlDeviceGetHandleByUUID(sid.encode())
            mx = p.nvmlDeviceGetMaxClockInfo(h, p.NVML_CLOCK_SM)
            bits = (getattr(p, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(p, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(p, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(p, "nvmlClocksEventReasonSwPowerCap", 0x4))
            for _ in range(2):
                sm = p.nvmlDeviceGetClockInfo(h, p.NVML_CLOCK_SM)
                try:
                    r = p.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = p.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits])
            self.how = "after"
        except Exception:
            pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": self.how}


# ------------------------------------------------------------------------------------------------------
class Rig:
    """One rank's share of a C3 scene of `n_total` sources: device-synthesised PCM (a private Frames block per source),
    a scene factory, and - for N > 1 - the peer-memory exchange."""

    def __init__(self, args, odb, torch, dist, ctx, stream, dev, rank, world, n_total, steps, warmup, spare=0):
        from oddio_b200.sharding import shard_sources

        self.args, self.odb, self.torch, self.dist = args, odb, torch, dist
        self.ctx, self.stream, self.dev, self.rank, self.world = ctx, stream, dev, rank, world
        self.N, self.M = n_total, args.frames
        self.pos, self.vel, freq, phase = scene_geometry(n_total)
        self.mine = shard_sources(n_total, rank, world)  # round-robin shard
        n_local = self.n_local = len(self.mine)
        M, K, W = self.M, steps, warmup
        # Every source plays its own PCM and every callback reads fresh samples, so the PCM resident in HBM grows with
        # K + W (19.6 GB for 65 536 sources and the default 16 + 3). If the requested number of steps does not fit this
        # GPU, the timed steps are cut to what fits and the JSON line says so - a shorter honest run instead of an
        # allocation failure.
        self.note = ""
        free_b, _total_b = torch.cuda.mem_get_info(dev)
        per_callback = int(np.ceil(DS_MAX * M)) * 4 * max(1, n_local)
        fit = int((0.8 * free_b - pcm_len(M, 0) * 4 * max(1, n_local)) // per_callback)
        if world > 1:  # every rank must time the same number of steps
            t_fit = torch.tensor([fit], device=dev, dtype=torch.int64)
            dist.all_reduce(t_fit, op=dist.ReduceOp.MIN)
            fit = int(t_fit.item())
        fit -= spare  # callbacks a pass may run on top of W + K (the exchange warm-up of the grouped modes)
        if K + W > fit:
            if fit - W < 1:
                raise SystemExit(f"bench: {n_local} sources x {M} frames do not fit this GPU even for one timed step")
            self.note = f"--steps {K} needs more PCM than fits in HBM; timed {fit - W} steps instead"
            K = fit - W
        self.K, self.W = K, W
        L = self.L = pcm_len(M, K + W + spare)
        t0 = time.time()
        self.frames = []
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x0DD10 + rank)
        kk = torch.arange(L, device=dev, dtype=torch.float32)
        B = 512
        for b0 in range(0, n_local, B):
            ids = self.mine[b0:b0 + B]
            w = torch.tensor(2 * np.pi * freq[ids] / RATE, device=dev, dtype=torch.float32)[:, None]
            ph = torch.tensor(phase[ids], device=dev, dtype=torch.float32)[:, None]
            x = 0.5 * torch.sin(w * kk[None, :] + ph) + 0.05 * (2 * torch.rand((len(ids), L), device=dev, generator=gen) - 1)
            x = x.contiguous()
            torch.cuda.synchronize(dev)
            for r in range(len(ids)):
                self.frames.append(odb.Frames.from_device(RATE, 1, x[r].data_ptr(), L, ctx))
            del x
        self.pcm_gb = n_local * L * 4 / 1e9
        self.setup_s = time.time() - t0
        self.interval = float(np.float32(1.0) / np.float32(RATE))  # oddio::run, lib.rs:91

    def new_scene(self):
        odb = self.odb
        ctl, scene = odb.SpatialScene.new(self.ctx)
        scene.set_kernel_variant(self.args.variant)
        handles = []
        for i, g in enumerate(self.mine):
            handles.append(ctl.play(odb.FramesSignal(self.frames[i], START_S), odb.SpatialOptions(self.pos[g], self.vel[g], 0.1)))
        return ctl, scene, handles

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def release(self):
        for f in self.frames:
            f.release()
        self.frames = []


def timed_device_pass(rig, clk, exchange_mode, reduce_every=1, depth=4, lag=1):
    """K callbacks back to back on the device (inputs resident in HBM), CUDA events on the launch stream.
    exchange_mode (N > 1): "kernel" = the exchange folded into the callback kernel (odb_scene_sample_exchange; every
    callback pushes its tile and receives the sum of callback k - lag), "peer" = stand-alone push / pull kernels once
    per `reduce_every` callbacks, "nccl" = torch.distributed all-reduce once per `reduce_every` callbacks."""
    torch, dist, world, dev, stream = rig.torch, rig.dist, rig.world, rig.dev, rig.stream
    M, K, W = rig.M, rig.K, rig.W
    from oddio_b200.sharding import PeerExchange

    note = ""
    with torch.cuda.stream(stream):
        ctl, scene, handles = rig.new_scene()
        if world == 1 or exchange_mode == "none":
            tiles = [torch.zeros((M, 2), device=dev, dtype=torch.float32) for _ in range(2)]
            k_no = [0]

            def step():
                scene.sample_device(rig.interval, tiles[k_no[0] & 1].data_ptr(), M)
                k_no[0] += 1

            def drain():
                pass

            def last_tile():
                return tiles[(k_no[0] - 1) & 1]
            launches_per_step = None
            exch = None
        elif exchange_mode == "kernel":
            exch = PeerExchange.from_torch(rig.ctx, M * 2, depth=max(2, min(8, depth)))
            tiles = [torch.zeros((M, 2), device=dev, dtype=torch.float32) for _ in range(lag + 2)]
            k_no, got = [0], [0]

            def step():
                if scene.sample_exchange(exch, rig.interval, tiles[got[0] % len(tiles)].data_ptr(), M, lag=lag):
                    got[0] += 1
                k_no[0] += 1

            def drain():
                while got[0] < k_no[0]:
                    exch.pull(tiles[got[0] % len(tiles)].data_ptr(), M * 2, 0, stream.cuda_stream)
                    got[0] += 1

            def last_tile():
                return tiles[(got[0] - 1) % len(tiles)]
            launches_per_step = None
        else:
            R = max(1, reduce_every)
            NG = max(2, min(8, depth))
            groups = [torch.zeros((R, M, 2), device=dev, dtype=torch.float32) for _ in range(NG)]
            comm = torch.cuda.Stream(device=dev, priority=-1)
            peer = exchange_mode == "peer"
            exch = None
            if peer:
                try:
                    exch = PeerExchange.from_torch(rig.ctx, R * M * 2, depth=NG)
                except rig.odb.OddioError as e:  # raised on every rank alike
                    peer, exch = False, None
                    note = f" (peer-memory exchange unavailable, fell back to NCCL: {str(e)[:120]})"
            pending = [False] * NG
            mixed = [torch.cuda.Event() for _ in range(NG)]
            reduced = [torch.cuda.Event() for _ in range(NG)]
            k_no = [0]

            def exchange(g):
                if peer:
                    exch.push(groups[g].data_ptr(), R * M * 2, comm.cuda_stream)
                    pending[g] = True
                else:
                    with torch.cuda.stream(comm):
                        dist.all_reduce(groups[g])

            def finish(g):
                if peer and pending[g]:
                    exch.pull(groups[g].data_ptr(), R * M * 2, 0, stream.cuda_stream)
                    pending[g] = False

            def step():
                k = k_no[0]
                k_no[0] += 1
                g, slot = (k // R) % NG, k % R
                if slot == 0:
                    stream.wait_event(reduced[g])
                    finish(g)
                scene.sample_device(rig.interval, groups[g][slot].data_ptr(), M)
                if slot == R - 1:
                    mixed[g].record(stream)
                    comm.wait_event(mixed[g])
                    exchange(g)
                    reduced[g].record(comm)

            def drain():
                k = k_no[0]
                if k % R:
                    g = (k // R) % NG
                    mixed[g].record(stream)
                    comm.wait_event(mixed[g])
                    exchange(g)
                    reduced[g].record(comm)
                    k_no[0] += R - k % R
                stream.wait_stream(comm)
                for i in range(NG):
                    finish((k_no[0] // R + i) % NG)

            def last_tile():
                return groups[((k_no[0] - 1) // R) % NG][(k_no[0] - 1) % R]
            launches_per_step = (2.0 / R) if peer else 0.0
        for _ in range(W):
            step()
        drain()
        if world > 1 and exchange_mode not in ("kernel", "none"):  # every group buffer / inbox slot has been through the exchange once
            for _ in range(NG * R):
                step()
            drain()
        own = scene.last_launch_count()
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with clk:
            e0.record(stream)
            h0 = time.perf_counter()
            for _ in range(K):
                step()
            drain()
            host_us = (time.perf_counter() - h0) / K * 1e6  # CPU time to queue one callback (the GPU runs behind)
            e1.record(stream)
            rig.barrier()
        ms = e0.elapsed_time(e1)
        counters = scene.last_job_counters()
        # the rare source with exactly one ear on FramesSignal's ds ~= 1 path takes the literal path
        assert counters["general"] <= max(1, rig.n_local // 1000), f"{counters['general']} jobs fell back to the literal path"
        assert scene.len() == rig.n_local, "a source finished during the timed region"
        checksum = float(last_tile().abs().sum().item())
        scene.close()
        if world > 1 and exchange_mode != "none" and exch is not None:
            exch.close()
    return {"ms": ms, "host_us": host_us, "counters": counters, "checksum": checksum, "note": note,
            "launches_per_step": own + (launches_per_step or 0.0)}


def kernel_time_pass(rig):
    """The dominant kernel's device time: a separate pass with CUDA events around the kernel (odb_set_profiling)."""
    torch, stream = rig.torch, rig.stream
    with torch.cuda.stream(stream):
        ctl, scene, handles = rig.new_scene()
        scene.set_profiling(True)
        tile = torch.zeros((rig.M, 2), device=rig.dev, dtype=torch.float32)
        kms = []
        for i in range(rig.W + rig.K):
            scene.sample_device(rig.interval, tile.data_ptr(), rig.M)
            if i >= rig.W:
                kms.append(scene.last_mix_kernel_ms())
        torch.cuda.synchronize(rig.dev)
        scene.close()
    return float(np.mean(kms))


def e2e_pass(rig, exchange_mode):
    """The same K callbacks through the calls a host makes, one at a time: the audio thread renders into HOST memory
    (one GPU: odb_scene_run with a host tile; N > 1: the in-kernel exchange with lag 0 writes the summed tile into
    pinned host memory) while a control thread calls set_motion on 1/16 of the sources per callback - the reference's
    two-thread architecture (README, examples/realtime.rs). H2D of the updates and D2H of the tile are inside the
    timed region; nothing is pipelined: callback k + 1 starts after the tile of callback k is in host memory."""
    torch, dist, odb, world, dev, stream = rig.torch, rig.dist, rig.odb, rig.world, rig.dev, rig.stream
    M, K, W = rig.M, rig.K, rig.W
    from oddio_b200.sharding import PeerExchange

    with torch.cuda.stream(stream):
        ctl, scene, handles = rig.new_scene()
        host_out = np.zeros((M, 2), dtype=np.float32)
        n_upd = max(1, rig.n_local // 16)
        ids_all = [h._src for h in handles]
        rng = np.random.default_rng(7 + rig.rank)
        upd = []
        for s in range(W + K):
            sel = rng.choice(rig.n_local, n_upd, replace=False)
            gl = rig.mine[sel]
            ids = (C.c_uint64 * n_upd)(*[ids_all[i] for i in sel])
            # the game thread nudges positions along the trajectory it already announced
            p = (rig.pos[gl] + rig.vel[gl] * np.float32((s + 1) * M / RATE)).astype(np.float32)
            upd.append((ids, p, rig.vel[gl].copy()))
        go = threading.Semaphore(0)

        def control_thread():
            for s in range(W + K):
                go.acquire()
                ids, p, v = upd[s]
                ctl.set_motion_ids(ids, n_upd, p, v)

        exch, pinned, dtile = None, None, None
        if world > 1:
            if exchange_mode == "nccl":
                dtile = torch.zeros((M, 2), device=dev, dtype=torch.float32)
            else:
                exch = PeerExchange.from_torch(rig.ctx, M * 2, depth=2)
                pinned = torch.zeros((M, 2), dtype=torch.float32, pin_memory=True)

        def step_e2e():
            go.release()
            if world == 1:
                odb.run(scene, RATE, host_out)
            elif exch is not None:  # the summed tile lands in pinned host memory straight from the kernel's reduce phase
                scene.sample_exchange(exch, rig.interval, pinned.data_ptr(), M, lag=0)
                rig.ctx.synchronize()
                host_out[:] = pinned.numpy()
            else:
                scene.sample_device(rig.interval, dtile.data_ptr(), M)
                dist.all_reduce(dtile)
                host_out[:] = dtile.cpu().numpy()

        th = threading.Thread(target=control_thread, daemon=True)
        th.start()
        for _ in range(W):
            step_e2e()
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        th.join()
        rig.barrier()
        e2e_s = time.perf_counter() - t0
        scene.close()
        if exch is not None:
            exch.close()
    return e2e_s, n_upd


def native_e2e(args, callbacks):
    """The e2e pass from compiled code: tools/e2e_native.c drives the C ABI exactly like e2e_pass (audio thread:
    odb_scene_run with a host tile; control thread: set_motion on 1/16 of the sources per callback) without an
    interpreter between the calls - what a Rust or C host sees. Built in-tree by __graft_entry__.build(); here only if
    the binary is missing and gcc is present. Returns its JSON line or None."""
    exe = os.path.join(ROOT, "tools", "e2e_native")
    if not os.path.exists(exe):
        try:
            import __graft_entry__ as ge

            ge.build_native_tools()
        except Exception:
            return None
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "oddio_b200") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env["ODB_VARIANT"] = str(args.variant)
    try:
        r = subprocess.run([exe, str(args.sources), str(callbacks), str(args.frames)], env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)
    except Exception:
        return None


def parity_pass(rig, n_sample=4096, callbacks=2):
    """N > 1, after the timed passes: a 4096-source scene sharded over the ranks like the timed one. Per callback,
    (i) the exchanged tile is bit-identical on every rank and equals the rank-order f32 sum of the per-rank tiles
    gathered with NCCL, (ii) rank 0 checks it against the CPU oracle's unsharded mix."""
    torch, dist, odb, world, dev, stream, rank = rig.torch, rig.dist, rig.odb, rig.world, rig.dev, rig.stream, rig.rank
    from oddio_b200.sharding import PeerExchange, shard_sources

    M = rig.M
    n = min(n_sample, rig.N)
    pos, vel, freq, phase = scene_geometry(rig.N)
    # trimmed private blocks: block i holds just what `callbacks` callbacks can touch (see tests/test_fullsize_gpu.py)
    dist_m = np.linalg.norm(pos[:n].astype(np.float64), axis=1)
    off = np.floor(RATE * (START_S - dist_m / 343.0)).astype(np.int64) - 128
    start = START_S - off.astype(np.float64) / RATE
    L = 128 + int(DS_MAX * M * callbacks) + 512
    kk = np.arange(L, dtype=np.float64)

    def block(i):
        r = np.random.default_rng(1000 + i)
        return (0.5 * np.sin(2 * np.pi * freq[i] / RATE * (kk + off[i]) + phase[i]) + 0.05 * r.uniform(-1, 1, L)).astype(np.float32)

    mine = shard_sources(n, rank, world)
    out = {"sources": n, "callbacks": callbacks}
    with torch.cuda.stream(stream):
        def shard_scene():
            ctl, sc = odb.SpatialScene.new(rig.ctx)
            sc.set_kernel_variant(rig.args.variant)
            for i in mine:
                ctl.play(odb.FramesSignal(odb.Frames.from_slice(RATE, block(i), rig.ctx), float(start[i])),
                         odb.SpatialOptions(pos[i], vel[i], 0.1))
            return ctl, sc
        _c1, plain = shard_scene()
        _c2, fused = shard_scene()
        exch = PeerExchange.from_torch(rig.ctx, M * 2, depth=2)
        tile = torch.zeros((M, 2), device=dev, dtype=torch.float32)
        xt = torch.zeros((M, 2), device=dev, dtype=torch.float32)
        ref_scene = None
        if rank == 0:
            from oracle import pyoracle as o

            ref_scene = o.SpatialScene()
            keep = []
            for i in range(n):
                fr = o.Frames.from_slice(RATE, block(i))
                keep.append(fr)
                ref_scene.play(o.FramesSignal(fr, float(start[i])), pos[i], vel[i], 0.1)
        ok_sum, ok_same, worst = True, True, 0.0
        for _ in range(callbacks):
            plain.sample_device(rig.interval, tile.data_ptr(), M)
            fused.sample_exchange(exch, rig.interval, xt.data_ptr(), M, lag=0)
            torch.cuda.synchronize(dev)
            parts = [torch.zeros_like(tile) for _ in range(world)]
            dist.all_gather(parts, tile)
            want = parts[0].clone()
            for r in range(1, world):
                want = want + parts[r]  # rank order, f32
            ok_sum = ok_sum and bool(torch.equal(want, xt))
            same = [torch.zeros_like(xt) for _ in range(world)]
            dist.all_gather(same, xt)
            ok_same = ok_same and all(bool(torch.equal(same[0], t)) for t in same)
            if rank == 0:
                ref = o.run(ref_scene, RATE, M).astype(np.float64)
                ref64 = ref_scene.out64(M)
                got = xt.cpu().numpy().astype(np.float64)
                rms = float(np.sqrt(np.mean(ref64 ** 2)))
                worst = max(worst, float(np.max(np.abs(got - ref) / (1e-5 * np.maximum(np.abs(ref), rms) + 1e-30))),
                            float(np.max(np.abs(got - ref64) / (1e-5 * np.maximum(np.abs(ref64), rms) + 1e-30))))
        flags = torch.tensor([int(ok_sum), int(ok_same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        plain.close(); fused.close(); exch.close()
    out.update({"exchanged_equals_rank_order_sum_bit_exact": bool(flags[0].item()), "identical_on_all_ranks": bool(flags[1].item()),
                "vs_oracle_worst_over_tolerance": worst, "tolerance": "1e-5 * max(|ref|, RMS) vs the reference-order f32 sum and the f64 truth",
                "vs_oracle_ok": worst <= 1.0})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import oddio_b200 as odb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = odb.Context(local, stream=stream.cuda_stream)  # our kernels and NCCL share one stream: no extra events
    clk = ClockSampler(local, enabled=(rank == 0)).start()  # initialises during the set-up below, far from the timed region
    M = args.frames
    mode = args.exchange if world > 1 else "none"

    def maxr(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the headline: the fixed scene of --sources sources split over the ranks (strong scaling), or --sources per rank
    N = args.sources * world if args.scaling == "weak" else args.sources
    spare = 0 if world == 1 else 8 * max(2, min(8, args.exchange_depth)) + 8
    rig = Rig(args, odb, torch, dist, ctx, stream, dev, rank, world, N, args.steps, args.warmup, spare)
    K, W = rig.K, rig.W
    main = timed_device_pass(rig, clk, mode, reduce_every=args.reduce_every, depth=args.exchange_depth, lag=args.lag)
    clk.stop()
    ms = maxr(main["ms"])
    kernel_ms = maxr(kernel_time_pass(rig))
    e2e_s, n_upd = (float("nan"), max(1, rig.n_local // 16)) if args.skip_e2e else e2e_pass(rig, mode)
    e2e_ms = maxr(e2e_s * 1e3)
    native = None
    if world == 1 and not args.skip_e2e and args.scaling == "strong":
        rig.release()  # its PCM makes room for the compiled harness's own scene
        native = native_e2e(args, K)
    extras = {}
    parity = None
    if world > 1 and not args.skip_extras:
        parity = parity_pass(rig)
        # the offline-rendering shape: stand-alone exchange of 8 callbacks at a time, overlapped with the next mixes
        off = timed_device_pass(rig, ClockSampler(local, enabled=False), "peer", reduce_every=8, depth=args.exchange_depth)
        extras["strong_reduce_every_8"] = {"value": N * M / (maxr(off["ms"]) / K * 1e-3), "unit": "source-frames/s",
                                           "note": "same scene; tiles summed by the stand-alone push / pull kernels once per 8 callbacks"}
        if args.scaling == "strong":
            rig.release()
            rig_w = Rig(args, odb, torch, dist, ctx, stream, dev, rank, world, args.sources * world, args.steps, args.warmup,
                        max(args.reduce_every, 1) * max(2, min(8, args.exchange_depth)) + 8)
            wk = timed_device_pass(rig_w, ClockSampler(local, enabled=False), mode, reduce_every=args.reduce_every,
                                   depth=args.exchange_depth, lag=args.lag)
            extras["weak"] = {"value": rig_w.N * M / (maxr(wk["ms"]) / rig_w.K * 1e-3), "unit": "source-frames/s",
                              "sources": rig_w.N, "sources_per_gpu": rig_w.n_local, "steps": rig_w.K,
                              "note": f"{args.sources} sources per GPU, exchange every callback (in the callback kernel)"}
            rig_w.release()

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: the PCM window, once per source per callback: 4 B * M * mean(ds)
        mine = rig.mine
        r = rig.pos[mine] / np.linalg.norm(rig.pos[mine], axis=1, keepdims=True)
        ds_mean = float(np.mean(1.0 - np.sum(rig.vel[mine] * r, axis=1) / 343.0))
        alg_bytes = 4.0 * M * ds_mean * rig.n_local
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        value = N * M / (ms / K * 1e-3)
        how = {"none": "", "kernel": f", tiles summed every callback from inside the callback kernel: its reduce phase stores the rank's sum "
                                     f"into every rank's inbox over NVLink peer memory ({M * 8} B per rank pair) and sums callback k - {args.lag} "
                                     "in rank order (odb_scene_sample_exchange)",
               "peer": f", tiles summed by the library's stand-alone peer-memory kernels over NVLink, one exchange per {args.reduce_every} callbacks, overlapped with the next mixes",
               "nccl": f", one NCCL all-reduce per {args.reduce_every} callbacks, overlapped with the next mixes",
               "none": "" if world == 1 else ", NO exchange of the tiles (diagnostic: the per-rank callback rate)"}[mode] + main["note"]
        legacy = bool(args.variant & 0x200)
        kname = ("k_mix_fast" if legacy else "k_scene_mix") + ("<strict>" if (args.variant & 0xFF) == 0 else "<fma>")
        out = {
            "metric": "source-frames/sec (N sources x buffer frames) spatial mix",
            "value": value, "unit": "source-frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C3 SpatialScene: {N} moving point sources (doppler + propagation delay), "
                                   f"{M}-frame stereo callback @{RATE} Hz, seek path (play)",
                       "sources": N, "sources_per_gpu": rig.n_local, "frames": M, "rate": RATE,
                       "parallelism": f"source-shard x{world}" + how,
                       "l2": "inputs larger than L2: every callback reads fresh PCM "
                             f"({alg_bytes / 1e6:.0f} MB per callback per GPU; {rig.pcm_gb:.1f} GB PCM resident per GPU)",
                       "kernel_variant": ("one-launch callback kernel" if not legacy else "round 1's multi-kernel callback") + (
                           ", strict (bit-exact per-source contributions)" if (args.variant & 0xFF) == 0 else
                           ", value multiply-adds contracted to FMA (cursors and indices bit-exact; the library default)"),
                       "jobs_last_callback": main["counters"], **({"note": rig.note} if rig.note else {}),
                       "host_enqueue_us_per_step": round(main["host_us"], 1), "setup_s": round(rig.setup_s, 1)},
            "clocks": clk.summary(),
            "e2e": {"value": native["source_frames_per_s"] if native else N * M / (e2e_ms / K * 1e-3), "unit": "source-frames/s",
                    "h2d_bytes_per_step": n_upd * 32, "d2h_bytes_per_step": M * 8 + 8,
                    "driver": ("compiled C harness over the C ABI (tools/e2e_native.c): what a Rust / C host sees" if native else
                               "Python (ctypes) over the C ABI"),
                    **({"us_per_callback": native["us_per_callback"], "python_driven_value": N * M / (e2e_ms / K * 1e-3)} if native else {}),
                    "note": ("audio thread: odb_scene_run with a host tile" if world == 1 else
                             "audio thread: odb_scene_sample_exchange (lag 0) on this rank's shard, the summed tile written to pinned host memory by the kernel"
                             if mode != "nccl" else "audio thread: sample_device + NCCL all-reduce + D2H copy")
                            + "; control thread: set_motion on 1/16 of the sources every callback"},
            "gpu_launches": int(round(main["launches_per_step"] * K)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("k_scene_mix" if not legacy else "k_mix_fast", rig.n_local), "kernel": kname,
                         "kernel_ms": kernel_ms, "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "kernel_share_of_step": kernel_ms / (ms / K),
                         "whole_callback_frac": alg_bytes / (ms / K * 1e-3) / 1e9 / peak},
            "checksum": main["checksum"],
        }
        if parity is not None:
            out["parity"] = parity
        if extras:
            out["extra"] = extras
        if args.cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, threads=1, n=args.cpu_sources)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------------
# BASELINE.json's other configurations (`--config`): parity-test cases first, but each with a reproducible number.
CONFIGS = {
    "C1": "Mixer: 8 Sine sources -> 1024 stereo frames @48 kHz on CPU (examples/offline.rs path); CPU only by SURVEY.md section 8",
    "C2": "SpatialScene: 1024 moving point sources, 256-frame callback @48 kHz",
    "C3": "SpatialScene: 65536 moving point sources, 1024-frame callback @48 kHz (the headline; default)",
    "C3b": "SpatialScene: the C3 sources through play_buffered(max_distance 350 m, 48 kHz, 0.1 s)",
    "C4": "Mixer: 262144 static stereo FramesSignal sources + Gain + Tanh, 1024 frames @96 kHz",
    "C5": "Mixer: 4096 Speed<FramesSignal> sources, ratio U[0.5, 2.0), 4096 frames @48 kHz",
}


def run_config(args):
    """One JSON line for C1 / C2 / C3b / C4 / C5 on one GPU: device-resident source-frames/s (CUDA events on the launch
    stream, L2 flushed between callbacks where the working set fits in it), the whole callback's algorithmic HBM
    fraction, the host-buffer call timed end to end, and the CPU oracle on a bounded sample."""
    from oracle import pyoracle as o

    cfg = args.config
    K, W = args.steps, max(3, args.warmup)
    if cfg == "C1":  # 8 x MonoToStereo(Sine) under a Mixer<[f32; 2]>, CPU only
        rng = np.random.default_rng(1)
        mx = o.Mixer(2)
        for _ in range(8):
            mx.play(o.MonoToStereo(o.Sine(float(rng.uniform(0, 2 * np.pi)), float(rng.uniform(100.0, 1000.0)))))
        reps = 2000
        secs = o.time_run(mx, RATE, 1024, 20, reps)  # best single callback
        line = {"metric": "source-frames/sec (N sources x buffer frames)", "value": 8 * 1024 / secs, "unit": "source-frames/s", "n_gpus": 0,
                "steps": reps, "warmup": 20, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": "C1 " + CONFIGS["C1"], "cpu_only": True},
                "gpu_launches": 0, "roofline": None,
                "cpu_baseline": {"value": 8 * 1024 / secs, "unit": "source-frames/s", "cores": 1, "kind": "port",
                                 "sample": f"the whole configuration, {reps} callbacks; C++ restatement of the Rust reference (no rustc here)"}}
        emit(line)
        return line
    import torch

    import oddio_b200 as odb

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    ctx = odb.Context(0, stream=stream.cuda_stream)
    clk = ClockSampler(0).start()
    rng = np.random.default_rng({"C2": 2, "C3b": 3, "C4": 4, "C5": 5}[cfg])
    scene_like = cfg in ("C2", "C3b")
    if cfg == "C2":
        N, M, rate, ch, ds_max = args.sources_cfg or 1024, 256, 48000, 1, DS_MAX
    elif cfg == "C3b":
        N, M, rate, ch, ds_max = args.sources_cfg or 65536, 1024, 48000, 1, 1.0
        W = max(W, 56)  # the delay rings fill for >= max_delay (1.12 s) before the output is non-trivial
    elif cfg == "C4":
        N, M, rate, ch, ds_max = args.sources_cfg or 262144, 1024, 96000, 2, 1.0
    else:
        N, M, rate, ch, ds_max = args.sources_cfg or 4096, 4096, 48000, 1, 2.0
    start_s = START_S if cfg == "C2" else 0.0
    L = int(start_s * rate + np.ceil(ds_max * M * (2 * (K + W) + 4))) + 2048   # device pass + e2e pass read disjoint PCM
    pos, vel, freq, phase = scene_geometry(N)
    if cfg == "C2":  # ball shell 2-100 m, velocities U[-30, 30] per axis
        d = pos / np.linalg.norm(pos, axis=1, keepdims=True)
        pos = (d * rng.uniform(2.0, 100.0, (N, 1))).astype(np.float32)
        vel = rng.uniform(-30, 30, (N, 3)).astype(np.float32)
    speeds = rng.uniform(0.5, 2.0, N).astype(np.float32) if cfg == "C5" else None
    gains = (rng.uniform(0.05, 1.0, N) * 1e-3).astype(np.float32) if cfg == "C4" else None
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    frames, host_pcm = [], {}
    n_cpu = {"C2": N, "C3b": 1024, "C4": 8192, "C5": 256}[cfg]
    kk = torch.arange(L, device=dev, dtype=torch.float32)
    t0 = time.time()
    B = 256
    for b0 in range(0, N, B):
        nb = min(B, N - b0)
        w = torch.tensor(2 * np.pi * freq[b0:b0 + nb] / rate, device=dev, dtype=torch.float32)[:, None, None]
        x = (0.5 * torch.sin(w * kk[None, :, None] + torch.arange(ch, device=dev)[None, None, :])
             + 0.05 * (2 * torch.rand((nb, L, ch), device=dev, generator=gen) - 1)).contiguous()
        torch.cuda.synchronize(dev)
        if b0 < n_cpu:
            xh = x[: max(0, min(nb, n_cpu - b0))].cpu().numpy()
            for r in range(xh.shape[0]):
                host_pcm[b0 + r] = xh[r, :, 0].copy() if ch == 1 else xh[r].copy()
        for r in range(nb):
            frames.append(odb.Frames.from_device(rate, ch, x[r].data_ptr(), L, ctx))
        del x
    setup_s = time.time() - t0

    def build(api, frames_of, n):
        """The configuration with `api` (the device mirror or the oracle) over its first n sources."""
        if scene_like:
            if api is odb:
                ctl, top = api.SpatialScene.new(ctx)
            else:
                top = api.SpatialScene()
                ctl = top
            for i in range(n):
                sig = api.FramesSignal(frames_of(i), start_s)
                if cfg == "C2":
                    ctl.play(sig, api.SpatialOptions(pos[i], vel[i], 0.1)) if api is odb else ctl.play(sig, pos[i], vel[i], 0.1)
                elif api is odb:
                    ctl.play_buffered(sig, api.SpatialOptions(pos[i], vel[i], 0.1), 350.0, rate, 0.1)
                else:
                    ctl.play_buffered(sig, pos[i], vel[i], 0.1, 350.0, rate, 0.1)
            return top, top
        if api is odb:
            ctl, mx = api.Mixer.new(ch, ctx)
        else:
            mx = api.Mixer(ch)
            ctl = mx
        top = api.Tanh(mx) if cfg == "C4" else mx
        for i in range(n):
            sig = api.FramesSignal(frames_of(i), 0.0)
            if cfg == "C4":
                if api is odb:
                    sig = api.Gain(sig)
                    sig.set_amplitude_ratio(float(gains[i]))
                else:
                    sig = api.Gain(sig)
                    sig.set_amplitude_ratio(float(gains[i]))
            else:
                if api is odb:
                    sc, sig = api.Speed.new(sig)
                    sc.set_speed(float(speeds[i]))
                else:
                    sig = api.Speed(sig)
                    sig.set_speed(float(speeds[i]))
            ctl.play(sig)
        return top, mx

    interval = float(np.float32(1.0) / np.float32(rate))
    flush = cfg in ("C2", "C5")  # working set per callback fits in the 126 MB L2: flush it between callbacks
    with torch.cuda.stream(stream):
        top, agg = build(odb, lambda i: frames[i], N)
        tile = torch.zeros((M, 2 if scene_like else ch), device=dev, dtype=torch.float32)
        scratch = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32) if flush else None  # 256 MB > L2
        for _ in range(W):
            top.sample_device(interval, tile.data_ptr(), M)
        torch.cuda.synchronize(dev)
        ms_list = []
        with clk:
            if flush:
                for _ in range(K):
                    scratch.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    top.sample_device(interval, tile.data_ptr(), M)
                    e1.record(stream)
                    torch.cuda.synchronize(dev)
                    ms_list.append(e0.elapsed_time(e1))
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(K):
                    top.sample_device(interval, tile.data_ptr(), M)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ms_list = [e0.elapsed_time(e1) / K] * K
        clk.stop()
        ms = float(np.mean(ms_list))
        launches = agg.last_launch_count()
        counters = agg.last_job_counters()
        checksum = float(tile.abs().sum().item())
        # end to end: the host-buffer call, one callback at a time
        host_out = np.zeros((M, 2 if scene_like else ch) if (scene_like or ch > 1) else (M,), dtype=np.float32)
        t1 = time.perf_counter()
        for _ in range(K):
            odb.run(top, rate, host_out)
        e2e_ms = (time.perf_counter() - t1) / K * 1e3
    # algorithmic bytes: 4 B * channels * M * ds per source and callback (+ 8 B per ring sample written for C3b)
    if cfg == "C2":
        r = pos / np.linalg.norm(pos, axis=1, keepdims=True)
        ds_mean = float(np.mean(1.0 - np.sum(vel * r, axis=1) / 343.0))
    else:
        ds_mean = float(speeds.mean()) if speeds is not None else 1.0
    alg = 4.0 * ch * M * ds_mean * N
    peak, peak_src = peaks()
    # CPU oracle on a bounded sample
    ofr = {}

    def oframes(i):
        if i not in ofr:
            ofr[i] = o.Frames.from_slice(rate, host_pcm[i])
        return ofr[i]
    otop, _ = build(o, oframes, n_cpu)
    reps = 3
    secs = o.time_run(otop, rate, M, 1, reps)  # best single callback
    line = {"metric": "source-frames/sec (N sources x buffer frames)", "value": N * M / (ms * 1e-3), "unit": "source-frames/s", "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{cfg} {CONFIGS[cfg]}", "sources": N, "frames": M, "rate": rate, "channels": ch,
                       "l2": "L2 flushed between callbacks (a 256 MB fill)" if flush else "inputs larger than L2: every callback reads fresh PCM",
                       "jobs_last_callback": counters, "setup_s": round(setup_s, 1)},
            "clocks": clk.summary(),
            "e2e": {"value": N * M / (e2e_ms * 1e-3), "unit": "source-frames/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(host_out.nbytes) + 8, "note": "odb_*_run with a host tile, Python (ctypes) over the C ABI"},
            "gpu_launches": int(launches * K),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "whole callback (all kernels of one *_sample)",
                         "kernel_ms": ms, "alg_bytes_per_launch": alg, "peak_source": peak_src},
            "checksum": checksum,
            "cpu_baseline": {"value": n_cpu * M / secs, "unit": "source-frames/s", "cores": 1, "kind": "port",
                             "sample": f"first {n_cpu} of the {N} sources x {M} frames, {reps} callbacks after 1 warm-up ({secs * 1e3:.1f} ms per "
                                       "callback); C++ restatement of the Rust reference (no rustc here)"}}
    emit(line)
    return line


# ------------------------------------------------------------------------------------------------------
def oracle_c3_scenes(args, n, shards, callbacks):
    """The first `n` sources of the C3 scene in the CPU oracle, dealt round-robin to `shards` scenes. Every source
    plays a private block trimmed to what `callbacks` callbacks can touch (block i starts off[i] frames into the sound
    and its FramesSignal starts at START_S - off[i] / rate, frames.rs:156), so the full 65 536-source scene is 6 GB of
    host memory instead of 19 GB and every callback still streams fresh PCM per source."""
    from oracle import pyoracle as o

    M = args.frames
    pos, vel, freq, phase = scene_geometry(args.sources)
    dist_m = np.linalg.norm(pos[:n].astype(np.float64), axis=1)
    off = np.floor(RATE * (START_S - dist_m / 343.0)).astype(np.int64) - 128
    start = START_S - off.astype(np.float64) / RATE
    L = 128 + int(DS_MAX * M * callbacks) + 512
    kk = np.arange(L, dtype=np.float32)
    rng = np.random.default_rng(1)
    scenes = [o.SpatialScene() for _ in range(shards)]
    keep = []
    B = 1024
    for b0 in range(0, n, B):
        ids = np.arange(b0, min(n, b0 + B))
        w = (2 * np.pi * freq[ids] / RATE).astype(np.float32)[:, None]
        ph = (phase[ids] + 2 * np.pi * freq[ids] / RATE * off[ids]).astype(np.float32)[:, None]
        x = (0.5 * np.sin(w * kk[None, :] + ph) + 0.05 * (2 * rng.random((len(ids), L), dtype=np.float32) - 1)).astype(np.float32)
        for r, i in enumerate(ids):
            fr = o.Frames.from_slice(RATE, x[r])
            keep.append(fr)
            scenes[i % shards].play(o.FramesSignal(fr, float(start[i])), pos[i], vel[i], 0.1)
    return scenes, keep


def cpu_baseline(args, threads: int, n: int, reps: int = 3, warm: int = 1):
    """Times the CPU oracle (oracle/, the restatement of the Rust reference) on the first `n` sources of the same
    scene, same frames per callback. threads = 1 is the reference's behaviour (one Signal graph, one thread,
    signal.rs:19); more threads shard the sources - not reference behaviour, labelled as such."""
    from oracle import pyoracle as o

    M = args.frames
    n = min(n, args.sources)
    scenes, keep = oracle_c3_scenes(args, n, threads, reps + warm)
    if threads == 1:
        t0 = time.perf_counter()
        for _ in range(warm):
            o.run(scenes[0], RATE, M)
        t1 = time.perf_counter()
        for _ in range(reps):
            o.run(scenes[0], RATE, M)
        per_cb = (time.perf_counter() - t1) / reps
        del t0
    else:
        secs, _tile = o.time_run_sharded(scenes, RATE, M, warm, reps)
        per_cb = secs / reps
    what = "all" if n == args.sources else "first"
    return {"value": n * M / per_cb, "unit": "source-frames/s", "cores": threads,
            "kind": "port", "sample": f"{what} {n} of the {args.sources} sources x {M} frames, {reps} callbacks after {warm} warm-up "
                                      f"({per_cb * 1e3:.1f} ms per callback); C++ restatement of the Rust reference (no rustc here)"
                                      + ("" if threads == 1 else f"; sources sharded over {threads} threads - not reference behaviour"),
            "same_config": n == args.sources, "host_cores_available": os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path - the oracle port (the Rust crate cannot be
    built here). The line's value is the reference's own behaviour: ONE thread over the FULL scene (65 536 sources,
    every step a whole callback); the all-host-threads run (sources sharded, not reference behaviour) is an extra key."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    M = args.frames
    N = args.sources  # the fixed scene; its CPU time does not depend on how many GPUs the other arm uses
    steps, warm = max(1, min(args.steps, args.ref_steps)), max(1, min(args.warmup, 3))
    base = cpu_baseline(args, threads=1, n=N, reps=steps, warm=warm)
    threads = os.cpu_count() or 1
    many = cpu_baseline(args, threads=threads, n=N, reps=steps, warm=warm) if threads > 1 else None
    line = {
        "impl": "reference", "metric": "source-frames/sec (N sources x buffer frames) spatial mix",
        "value": base["value"], "unit": "source-frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": N * M / base["value"] * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C3 SpatialScene: {N} moving point sources (doppler + propagation delay), "
                               f"{M}-frame stereo callback @{RATE} Hz, seek path (play)",
                   "sources": N, "frames": M, "rate": RATE, "parallelism": "1 host thread (the reference's Signal graph is single-threaded)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "source-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.time() - t0, 1),
    }
    if many:
        line["extra"] = {"all_host_threads": many}
    emit(line)


def main():
    # Exactly one JSON line may reach stdout: libraries that print there (NCCL's version banner) are sent to
    # stderr by pointing fd 1 at fd 2 for the whole run; the result line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="; ".join(f"{k}: {v}" for k, v in sorted(CONFIGS.items())))
    ap.add_argument("--sources-cfg", type=int, default=0, help="--config other than C3: source count (0 = the configuration's own)")
    ap.add_argument("--sources", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--cpu-sources", type=int, default=2048, help="sources in the bounded CPU sample")
    ap.add_argument("--ref-steps", type=int, default=32, help="--impl reference: cap on the timed whole-scene callbacks (about 0.85 s each on one thread)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the fixed scene of --sources sources split over the ranks, weak = --sources per GPU")
    ap.add_argument("--lag", type=int, default=1, help="N > 1, --exchange kernel: callback k receives the summed tile of callback k - lag")
    ap.add_argument("--skip-extras", action="store_true", help="N > 1: headline only (no parity pass, no weak-scaling / reduce-every-8 passes)")
    ap.add_argument("--reduce-every", type=int, default=1, help="N > 1, --exchange peer|nccl: callbacks per exchange of the tiles")
    ap.add_argument("--exchange-depth", type=int, default=4, help="N > 1: group buffers (exchanges) in flight, 2..8")
    ap.add_argument("--exchange", default="kernel", choices=["kernel", "peer", "nccl", "none"],
                    help="N > 1: how the per-GPU tiles are summed: kernel (default) = from inside the callback kernel over NVLink peer "
                         "memory, every callback; peer = the library's stand-alone push / pull kernels; nccl = torch.distributed")
    ap.add_argument("--variant", type=lambda v: int(v, 0), default=2,
                    help="2 (default, also the library's) = value ops contracted to FMA, 0 = strict (bit-exact per-source "
                         "contributions); | 0x200 = round 1's multi-kernel callback instead of the one-launch kernel")
    ap.add_argument("--skip-e2e", action="store_true", help="kernel experiments: device-resident passes only")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.config != "C3":
        if int(os.environ.get("RANK", "0")) == 0:
            run_config(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
