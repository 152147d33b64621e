#!/usr/bin/env python
"""bench.py — source-frames/s of the spatial-mix hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3"): one SpatialScene with 65 536 moving point sources
(`SpatialSceneControl::play` of a FramesSignal each: propagation delay, doppler resampling, distance
attenuation, stereo pan), 1024 stereo frames per callback at 48 kHz. A "step" is one `oddio::run`
callback over all sources. Every source owns its own PCM (no sharing; 19 GB in HBM in total), every
callback reads fresh PCM, so the inputs of a step are larger than L2 by construction.

One JSON line on stdout (rank 0). `value` = N_sources * frames / device time per callback with
everything resident in HBM; `e2e` = the same through the host-buffer C-ABI call
(`odb_scene_run`: host output tile, D2H inside the timed region, plus `set_motion` updates for 1/16 of
the sources every callback, H2D inside the timed region); `roofline` = algorithmic PCM bytes of the
staged mix kernel / its device time against the measured HBM peak; `cpu_baseline` = the CPU oracle
(the only runnable statement of the Rust reference here) on a bounded sample of the same workload.

N > 1: one process per GPU; every rank mixes its own shard of the scene's sources and each callback ends
with a sum all-reduce of the 8 KiB stereo tile. The sum is linear, so `--reduce-every R` (default 8, the offline
rendering shape of examples/offline.rs) exchanges R tiles in one NCCL all-reduce issued on a second stream that
overlaps the next callbacks' mixes; the timed region ends after the last all-reduce. R = 1 is live playback. Default `--scaling weak`: 65 536 sources per GPU (the scene grows with
the box); `--scaling strong`: the 65 536 sources are split over the ranks.
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

    from oddio_b200.sharding import shard_sources

    M, K, W = args.frames, args.steps, args.warmup
    N = args.sources * world if args.scaling == "weak" else args.sources  # sources of the whole job
    pos, vel, freq, phase = scene_geometry(N)
    mine = shard_sources(N, rank, world)  # round-robin shard
    n_local = len(mine)
    # Every source plays its own PCM and every callback reads fresh samples, so the PCM resident in HBM grows with
    # K + W (19.6 GB for the default 16 + 3). If the requested number of steps does not fit this GPU, the timed
    # steps are cut to what fits and the JSON line says so in "steps" and config.note - a shorter honest run
    # instead of an allocation failure.
    steps_note = ""
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    per_callback = int(np.ceil(DS_MAX * M)) * 4 * max(1, n_local)
    fit = int((0.8 * free_b - pcm_len(M, 0) * 4 * max(1, n_local)) // per_callback)
    if world > 1:  # every rank must time the same number of steps
        t_fit = torch.tensor([fit], device=dev, dtype=torch.int64)
        dist.all_reduce(t_fit, op=dist.ReduceOp.MIN)
        fit = int(t_fit.item())
    if K + W > fit:
        if fit - W < 1:
            raise SystemExit(f"bench: {n_local} sources x {M} frames do not fit this GPU even for one timed step")
        steps_note = f"--steps {K} needs more PCM than fits in HBM; timed {fit - W} steps instead"
        K = fit - W
    L = pcm_len(M, K + W)

    # ---- synthetic PCM, generated on the device, one private Frames block per source ------------------
    t_setup = time.time()
    frames = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x0DD10 + rank)
    kk = torch.arange(L, device=dev, dtype=torch.float32)
    B = 512
    for b0 in range(0, n_local, B):
        ids = mine[b0:b0 + B]
        w = torch.tensor(2 * np.pi * freq[ids] / RATE, device=dev, dtype=torch.float32)[:, None]
        ph = torch.tensor(phase[ids], device=dev, dtype=torch.float32)[:, None]
        x = 0.5 * torch.sin(w * kk[None, :] + ph) + 0.05 * (2 * torch.rand((len(ids), L), device=dev, generator=gen) - 1)
        x = x.contiguous()
        torch.cuda.synchronize(dev)
        for r in range(len(ids)):
            frames.append(odb.Frames.from_device(RATE, 1, x[r].data_ptr(), L, ctx))
        del x
    pcm_gb = n_local * L * 4 / 1e9

    def new_scene():
        ctl, scene = odb.SpatialScene.new(ctx)
        scene.set_kernel_variant(args.variant)
        handles = []
        for i, g in enumerate(mine):
            handles.append(ctl.play(odb.FramesSignal(frames[i], START_S), odb.SpatialOptions(pos[g], vel[g], 0.1)))
        return ctl, scene, handles

    interval = float(np.float32(1.0) / np.float32(RATE))  # oddio::run, lib.rs:91
    # Offline rendering batches R consecutive callbacks per exchange: the sum over ranks is linear, so one
    # all-reduce of R tiles equals R all-reduces of one tile. R = 1 is the live-playback shape.
    # --exchange peer (default): the library's own one-kernel push/sum over NVLink peer memory; --exchange nccl:
    # torch.distributed all-reduce. Either way R tiles per exchange (--reduce-every; 1 = the live-playback shape).
    peer = world > 1 and args.exchange == "peer"
    R = 1 if world == 1 else max(1, args.reduce_every)
    # NG group buffers in flight: a rank may run up to NG - 1 exchanges ahead of the slowest one, which absorbs the
    # host-side jitter of the other ranks instead of paying max-over-ranks at every exchange
    NG = max(2, min(8, args.exchange_depth)) if world > 1 else 2
    groups = [torch.zeros((R, M, 2), device=dev, dtype=torch.float32) for _ in range(NG)]
    comm = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    exch = None
    exchange_note = ""
    if peer:
        from oddio_b200.sharding import PeerExchange

        try:
            exch = PeerExchange.from_torch(ctx, R * M * 2, depth=NG)
        except odb.OddioError as e:  # raised on every rank alike (e.g. CUDA IPC not permitted in this container)
            peer, exch = False, None
            exchange_note = f" (peer-memory exchange unavailable, fell back: {str(e)[:120]})"

    pending = [False] * NG  # peer exchange: group g has been pushed and not yet pulled

    def exchange(g):
        """Sum of group g over the ranks, started on the `comm` stream. Peer exchange: only the push half (it never
        waits for another rank); the pull half is queued one group later, right before group g's buffer is reused,
        when every rank has long pushed - a pipelined renderer consumes the summed tiles one group late."""
        if peer:
            exch.push(groups[g].data_ptr(), R * M * 2, comm.cuda_stream)
            pending[g] = True
        else:
            with torch.cuda.stream(comm):
                dist.all_reduce(groups[g])

    def finish_exchange(g):
        if peer and pending[g]:
            exch.pull(groups[g].data_ptr(), R * M * 2, 0, stream.cuda_stream)
            pending[g] = False
    mixed = [torch.cuda.Event() for _ in range(NG)]
    reduced = [torch.cuda.Event() for _ in range(NG)]
    step_no = [0]

    def step_device(scene):
        """One callback: mix this rank's shard into slot k%R of group (k/R)%2 on `stream`; after R callbacks
        the NCCL sum of the group runs on `comm` and overlaps the next group's mixes."""
        k = step_no[0]
        step_no[0] += 1
        g, slot = (k // R) % NG, k % R
        if world > 1 and slot == 0:
            stream.wait_event(reduced[g])  # the group buffer is free again once its previous exchange has left
            finish_exchange(g)             # (peer exchange: the sum of the group's previous contents lands here)
        scene.sample_device(interval, groups[g][slot].data_ptr(), M)
        if world > 1 and slot == R - 1:
            mixed[g].record(stream)
            comm.wait_event(mixed[g])
            exchange(g)
            reduced[g].record(comm)

    def drain():
        if world > 1:
            k = step_no[0]
            if k % R:  # flush a partial group
                g = (k // R) % NG
                mixed[g].record(stream)
                comm.wait_event(mixed[g])
                exchange(g)
                reduced[g].record(comm)
                step_no[0] += R - k % R
            stream.wait_stream(comm)
            for i in range(NG):  # oldest first: pulls follow push order
                finish_exchange((step_no[0] // R + i) % NG)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- (1) device-resident throughput ------------------------------------------------------------------
    with torch.cuda.stream(stream):
        ctl, scene, handles = new_scene()
        setup_s = time.time() - t_setup
        for _ in range(W):
            step_device(scene)
        drain()
        if world > 1:  # both buffers / inbox parities of the exchange have been through it once before the clock starts
            first = (step_no[0] // R) % NG  # keep the rotation: the timed loop continues with this group
            for i in range(NG):
                g = (first + i) % NG
                mixed[g].record(stream)
                comm.wait_event(mixed[g])
                exchange(g)
                reduced[g].record(comm)
            stream.wait_stream(comm)
            for i in range(NG):
                finish_exchange((first + i) % NG)
        # our kernels per callback: the scene's own, plus the exchange's push and pull once per R callbacks (an NCCL
        # all-reduce is not ours and is not counted)
        launches_per_step = scene.last_launch_count() + ((2.0 / R) if peer else 0.0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with clk:
            e0.record(stream)
            h0 = time.perf_counter()
            for _ in range(K):
                step_device(scene)
            drain()
            host_enqueue_us = (time.perf_counter() - h0) / K * 1e6  # CPU time to queue one callback (GPU runs behind)
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        counters = scene.last_job_counters()
        # the rare source with exactly one ear on FramesSignal's ds ~= 1 path takes the literal kernel
        assert counters["general"] <= max(1, n_local // 1000), f"{counters['general']} jobs fell back to the general kernel"
        assert scene.len() == n_local, "a source finished during the timed region"
        checksum = float(groups[((step_no[0] - 1) // R) % NG][(step_no[0] - 1) % R].abs().sum().item())
        clk.stop()
        scene.close()

        # ---- (2) mix-kernel device time for the roofline (separate pass: events around the kernel) ------------
        ctl, scene, handles = new_scene()
        scene.set_profiling(True)
        kms = []
        for i in range(W + K):
            step_device(scene)
            if i >= W:
                kms.append(scene.last_mix_kernel_ms())
        drain()
        scene.close()

        # ---- (3) end to end through the host-buffer call --------------------------------------------------------
        def run_e2e():
            ctl, scene, handles = new_scene()
            host_out = np.zeros((M, 2), dtype=np.float32)
            n_upd = max(1, n_local // 16)
            ids_all = [h._src for h in handles]
            rng = np.random.default_rng(7 + rank)
            upd = []
            for s in range(W + K):
                sel = rng.choice(n_local, n_upd, replace=False)
                gl = mine[sel]
                ids = (C.c_uint64 * n_upd)(*[ids_all[i] for i in sel])
                # the game thread nudges positions along the trajectory it already announced
                p = (pos[gl] + vel[gl] * np.float32((s + 1) * M / RATE)).astype(np.float32)
                upd.append((ids, p, vel[gl].copy()))

            # The reference's architecture has two threads: a control ("game") thread that calls set_motion and the
            # audio thread that calls run (README, examples/realtime.rs). Same here: the control thread queues the
            # updates of callback s while the audio thread is inside odb_scene_run of callback s (ctypes releases the GIL).
            # The audio thread paces the control thread with a semaphore (one batch of updates per callback) and never
            # waits for it.
            go = threading.Semaphore(0)

            def control_thread():
                for s in range(W + K):
                    go.acquire()
                    ids, p, v = upd[s]
                    ctl.set_motion_ids(ids, n_upd, p, v)

            e2e_tile = torch.zeros((M, 2), device=dev, dtype=torch.float32)

            def step_e2e(s):
                go.release()
                if world == 1:
                    odb.run(scene, RATE, host_out)  # host tile: H2D of the queued updates and D2H of the result inside
                else:  # this rank's shard into a device tile, summed over the ranks, then read back
                    scene.sample_device(interval, e2e_tile.data_ptr(), M)
                    if peer:
                        exch.allreduce(e2e_tile.data_ptr(), M * 2, 0, stream.cuda_stream)
                    else:
                        dist.all_reduce(e2e_tile)
                    host_out[:] = e2e_tile.cpu().numpy()

            th = threading.Thread(target=control_thread, daemon=True)
            th.start()
            for s in range(W):
                step_e2e(s)
            barrier()
            t0 = time.perf_counter()
            for s in range(W, W + K):
                step_e2e(s)
            th.join()
            barrier()
            e2e_s = time.perf_counter() - t0
            scene.close()
            return e2e_s


        e2e_s = float("nan") if args.skip_e2e else run_e2e()
        n_upd = max(1, n_local // 16)

    # max over ranks
    times = torch.tensor([ms, e2e_s * 1e3, float(np.mean(kms))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = [float(x) for x in times.tolist()]

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: the PCM window, once per source per callback: 4 B * M * mean(ds)
        r = pos[mine] / np.linalg.norm(pos[mine], axis=1, keepdims=True)
        ds_mean = float(np.mean(1.0 - np.sum(vel[mine] * r, axis=1) / 343.0))
        alg_bytes = 4.0 * M * ds_mean * n_local
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        value = N * M / (ms / K * 1e-3)
        out = {
            "metric": "source-frames/sec (N sources x buffer frames) spatial mix",
            "value": value, "unit": "source-frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C3 SpatialScene: {N} moving point sources (doppler + propagation delay), "
                                   f"{M}-frame stereo callback @{RATE} Hz, seek path (play)",
                       "sources": N, "sources_per_gpu": n_local, "frames": M, "rate": RATE,
                       "parallelism": f"source-shard x{world}" + ("" if world == 1 else (
                           f", tiles summed by the library's peer-memory kernel over NVLink, one exchange per {R} callbacks ({R * M * 8} B per rank pair), overlapped with the next mixes"
                           if peer else f", one NCCL all-reduce per {R} callbacks ({R * M * 8} B), overlapped with the next mixes{exchange_note}")),
                       "l2": "inputs larger than L2: every callback reads fresh PCM "
                             f"({alg_bytes / 1e6:.0f} MB per callback per GPU; {pcm_gb:.1f} GB PCM resident per GPU)",
                       "kernel_variant": ("staged, strict (bit-exact per-source contributions)" if args.variant == 0 else
                                          "staged, value multiply-adds contracted to FMA (cursors and indices bit-exact)"),
                       "jobs_last_callback": counters, **({"note": steps_note} if steps_note else {}), "host_enqueue_us_per_step": round(host_enqueue_us, 1),
                       "setup_s": round(setup_s, 1)},
            "clocks": clk.summary(),
            "e2e": {"value": N * M / (e2e_ms / K * 1e-3), "unit": "source-frames/s",
                    "h2d_bytes_per_step": n_upd * 32, "d2h_bytes_per_step": M * 8 + 4,
                    "note": ("audio thread: odb_scene_run with a host tile" if world == 1 else
                             "audio thread: odb_scene_sample_device on this rank's shard, the tiles summed over the ranks every callback (push + pull), read back to the host")
                            + "; control thread: set_motion on 1/16 of the sources every callback"},
            "gpu_launches": int(round(launches_per_step * K)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("k_mix_fast", n_local), "kernel": "k_mix_fast<strict>" if args.variant == 0 else "k_mix_fast<fma>", "kernel_ms": kernel_ms,
                         "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "kernel_share_of_step": kernel_ms / (ms / K)},
            "checksum": checksum,
        }
        if args.cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, threads=1, n=args.cpu_sources)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------------
def cpu_baseline(args, threads: int, n: int, reps: int = 3, warm: int = 1):
    """Times the CPU oracle (oracle/, the restatement of the Rust reference) on a bounded sample of the
    same workload: the first `n` sources of the same scene, same frames per callback."""
    from oracle import pyoracle as o

    M = args.frames
    n = min(n, args.sources)
    pos, vel, freq, phase = scene_geometry(args.sources)
    L = pcm_len(M, reps + warm)
    rng = np.random.default_rng(1)
    kk = np.arange(L, dtype=np.float32)
    scenes = [o.SpatialScene() for _ in range(threads)]
    keep = []
    for i in range(n):
        x = (0.5 * np.sin(np.float32(2 * np.pi * freq[i] / RATE) * kk + np.float32(phase[i]))
             + 0.05 * (2 * rng.random(L, dtype=np.float32) - 1)).astype(np.float32)
        fr = o.Frames.from_slice(RATE, x)
        keep.append(fr)
        scenes[i % threads].play(o.FramesSignal(fr, START_S), pos[i], vel[i], 0.1)
    secs, tile = o.time_run_sharded(scenes, RATE, M, warm, reps)
    per_cb = secs / reps
    return {"value": n * M / per_cb, "unit": "source-frames/s", "cores": threads,
            "kind": "port", "sample": f"first {n} of the {args.sources} sources x {M} frames, {reps} callbacks after {warm} warm-up "
                                      f"({per_cb * 1e3:.1f} ms per callback); C++ restatement of the Rust reference (no rustc here)",
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port; the Rust crate
    cannot be built here) with all host threads, sources sharded over threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.time()
    base = cpu_baseline(args, threads=threads, n=args.ref_sources, reps=args.steps, warm=args.warmup)
    M = args.frames
    N = args.sources * args.gpus if args.scaling == "weak" else args.sources
    line = {
        "impl": "reference", "metric": "source-frames/sec (N sources x buffer frames) spatial mix",
        "value": base["value"], "unit": "source-frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": N * M / base["value"] * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C3 SpatialScene: {N} moving point sources (doppler + propagation delay), "
                               f"{M}-frame stereo callback @{RATE} Hz, seek path (play)",
                   "sources": N, "frames": M, "rate": RATE, "parallelism": f"{threads} host threads, sources sharded"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "source-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.time() - t0, 1),
    }
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
    ap.add_argument("--sources", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--cpu-sources", type=int, default=2048, help="sources in the bounded CPU sample")
    ap.add_argument("--ref-sources", type=int, default=8192, help="sources in the --impl reference sample")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = --sources per GPU (default), strong = --sources in total")
    ap.add_argument("--reduce-every", type=int, default=8, help="N > 1: callbacks per exchange of the tiles (1 = live playback)")
    ap.add_argument("--exchange-depth", type=int, default=4, help="N > 1: group buffers (exchanges) in flight, 2..8")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the per-GPU tiles are summed (peer = the library's NVLink peer-memory kernel, every callback)")
    ap.add_argument("--variant", type=lambda v: int(v, 0), default=2,
                    help="2 (default, also the library's) = value ops contracted to FMA, 0 = strict (bit-exact per-source "
                         "contributions); | 0x200 = round 1's multi-kernel callback instead of the one-launch kernel")
    ap.add_argument("--skip-e2e", action="store_true", help="kernel experiments: device-resident passes only")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
