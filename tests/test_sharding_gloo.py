"""world_size-2 gloo test of the multi-rank logic: round-robin source sharding + one sum
all-reduce of the stereo tile reproduces the unsharded mix. Runs on CPU; the per-shard tiles come from the
oracle (the checker), since there is no GPU here - the GPU side of the same logic is bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_src, n_frames, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from helpers import rand_in_shell, synth_pcm
    from oddio_b200.sharding import allreduce_tile, shard_sources
    from oracle import pyoracle as o

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # same scene description on every rank
    pcms = [synth_pcm(rng, 60000, 48000) for _ in range(4)]
    pos = [rand_in_shell(rng, 2, 100) for _ in range(n_src)]
    vel = [rng.uniform(-30, 30, 3).astype(np.float32) for _ in range(n_src)]
    mine = shard_sources(n_src, rank, world)
    scene = o.SpatialScene()
    frames = [o.Frames.from_slice(48000, p) for p in pcms]
    for s in mine:
        scene.play(o.FramesSignal(frames[s % 4], 1.0), pos[s], vel[s], 0.1)
    outs = []
    for _ in range(3):
        tile = torch.from_numpy(o.run(scene, 48000, n_frames))
        allreduce_tile(tile)
        outs.append(tile.numpy().copy())
    if rank == 0:
        full = o.SpatialScene()
        for s in range(n_src):
            full.play(o.FramesSignal(frames[s % 4], 1.0), pos[s], vel[s], 0.1)
        refs = [o.run(full, 48000, n_frames) for _ in range(3)]
        q.put((outs, refs, [int(x) for x in mine]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_sources_partition():
    from oddio_b200.sharding import shard_sources

    for n, w in ((65536, 8), (10, 4), (3, 8), (0, 2)):
        parts = [shard_sources(n, r, w) for r in range(w)]
        allv = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allv, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_sources(4, 2, 2)


def test_two_rank_shard_and_allreduce_matches_unsharded():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_src, n_frames = 37, 512
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_src, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs, refs, mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mine == list(range(0, n_src, 2))
    for out, ref in zip(outs, refs):
        rms = float(np.sqrt(np.mean(ref.astype(np.float64) ** 2)))
        assert np.all(np.abs(out.astype(np.float64) - ref) <= 1e-5 * np.maximum(np.abs(ref), rms))
