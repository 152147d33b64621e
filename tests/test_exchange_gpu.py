"""The multi-GPU tile exchange over NVLink peer memory (csrc/odb_exchange.cu) through the C ABI.

* one rank: the exchange is the identity plus the epilogue (every code path of the kernel but the remote stores);
* 2 / 4 / 8 ranks (each needs a box with that many GPUs, skipped otherwise): one process per GPU, shard a SpatialScene
  round-robin, mix their shards with the CUDA path and sum the tiles with the exchange; rank 0 checks the result
  against the CPU oracle's unsharded mix on the same inputs, callback after callback (both inbox parities),
  and that both ranks hold bit-identical tiles.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_single_rank_exchange_is_identity_plus_epilogue():
    import torch

    import oddio_b200 as odb
    from oddio_b200.sharding import PeerExchange

    ctx = odb.init(0)
    ex = PeerExchange(ctx, 0, 1, 4096, depth=2)
    rng = np.random.default_rng(3)
    for n in (2048, 514, 2, 4096, 3000):  # 16-byte multiples, tails, several slices
        x = rng.uniform(-2, 2, n).astype(np.float32)
        for epi in (0, 1, 2):
            t = torch.from_numpy(x).cuda()
            ex.allreduce(t.data_ptr(), n, epi, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = t.cpu().numpy()
            if epi == 0:
                assert np.array_equal(got, x)
            elif epi == 1:
                np.testing.assert_allclose(got, np.tanh(x.astype(np.float64)), rtol=0, atol=4e-7)  # tanh.rs:26
            else:
                assert np.array_equal(got, x / (np.float32(1.0) + np.abs(x)))                       # reinhard.rs:31
    with pytest.raises(odb.OddioError):
        ex.allreduce(0, 8, 0)          # NULL tile
    t = torch.zeros(8192, device="cuda")
    with pytest.raises(odb.OddioError):
        ex.allreduce(t.data_ptr(), 8192, 0)  # beyond the capacity given at creation
    ex.close()


def _worker(rank, world, port, n_src, n_frames, n_callbacks, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oddio_b200 as odb
    from helpers import rand_in_shell, synth_pcm
    from oddio_b200.sharding import PeerExchange, shard_sources

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)  # set-up handles only
    ctx = odb.Context(rank)
    ex = PeerExchange.from_torch(ctx, 2 * n_frames, depth=3)
    rng = np.random.default_rng(11)  # the same scene description on every rank
    pcms = [synth_pcm(rng, 60000, 48000) for _ in range(4)]
    pos = [rand_in_shell(rng, 2, 100) for _ in range(n_src)]
    vel = [rng.uniform(-30, 30, 3).astype(np.float32) for _ in range(n_src)]
    mine = shard_sources(n_src, rank, world)
    frames = [odb.Frames.from_slice(48000, p, ctx) for p in pcms]

    def shard_scene():
        ctl, scene = odb.SpatialScene.new(ctx)
        for s in mine:
            ctl.play(odb.FramesSignal(frames[s % 4], 1.0), odb.SpatialOptions(pos[s], vel[s], 0.1))
        return ctl, scene

    interval = float(np.float32(1.0) / np.float32(48000))
    tile = torch.zeros((n_frames, 2), device=f"cuda:{rank}", dtype=torch.float32)
    # (a) the callback renders this rank's tile, the stand-alone kernels exchange it
    ctl, scene = shard_scene()
    outs = []
    for _ in range(n_callbacks):
        scene.sample_device(interval, tile.data_ptr(), n_frames)
        ex.allreduce(tile.data_ptr(), 2 * n_frames, 0)
        ctx.synchronize()
        outs.append(tile.cpu().numpy().copy())
    # (b) the exchange folded into the callback kernel, live (lag 0) and pipelined (lag 1, the last tile pulled by the
    # stand-alone kernel): the same per-rank sums in the same rank order, so bit-identical to (a)
    for lag in (0, 1):
        ex2 = PeerExchange.from_torch(ctx, 2 * n_frames, depth=3)
        ctl2, scene2 = shard_scene()
        got = []
        for k in range(n_callbacks):
            tile.fill_(-7.0)
            wrote = scene2.sample_exchange(ex2, interval, tile.data_ptr(), n_frames, lag=lag)
            ctx.synchronize()
            assert wrote == (k >= lag)
            if wrote:
                got.append(tile.cpu().numpy().copy())
        for _ in range(lag):
            ex2.pull(tile.data_ptr(), 2 * n_frames)
            ctx.synchronize()
            got.append(tile.cpu().numpy().copy())
        assert len(got) == n_callbacks
        for k in range(n_callbacks):
            assert np.array_equal(got[k], outs[k]), f"rank {rank}: in-kernel exchange (lag {lag}) differs at callback {k}"
        dist.barrier()
        scene2.close()
        ex2.close()
    # the pipelined form: `depth` pushes in flight, pulls late and into other buffers, slots reused several times
    nf = 2 * n_frames
    for rep in range(4):
        src = [torch.full((nf,), float((i + 1) * (rank + 1) + rep), device=f"cuda:{rank}") for i in range(3)]
        dst = [torch.empty_like(x) for x in src]
        for x in src:
            ex.push(x.data_ptr(), nf)
        for y in dst:
            ex.pull(y.data_ptr(), nf)
        ctx.synchronize()
        for i, y in enumerate(dst):
            assert float(y.min()) == float(y.max()) == sum((i + 1) * (r + 1) + rep for r in range(world))
    try:
        for _ in range(4):
            ex.push(src[0].data_ptr(), nf)
        raise AssertionError("a fourth outstanding push must be refused at depth 3")
    except odb.OddioError:
        for _ in range(3):
            ex.pull(dst[0].data_ptr(), nf)
        ctx.synchronize()
    q.put((rank, outs))
    dist.barrier()
    scene.close()
    ex.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_scene_matches_the_oracle(oracle, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    import torch.multiprocessing as mp

    from helpers import assert_mix_close, rand_in_shell, synth_pcm

    n_src, n_frames, n_cb = 37 + 16 * world, 1024, 4
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, n_src, n_frames, n_cb, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time

    got, deadline = {}, time.time() + 300
    while len(got) < world:  # fail fast if a worker died instead of waiting for the queue
        try:
            r, outs = q.get(timeout=1.0)
            got[r] = outs
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            assert not dead and time.time() < deadline, f"worker(s) failed: exit codes {[p.exitcode for p in procs]}"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the unsharded mix on the CPU oracle, same inputs (same seed as the workers)
    o = oracle
    rng = np.random.default_rng(11)
    pcms = [synth_pcm(rng, 60000, 48000) for _ in range(4)]
    pos = [rand_in_shell(rng, 2, 100) for _ in range(n_src)]
    vel = [rng.uniform(-30, 30, 3).astype(np.float32) for _ in range(n_src)]
    full = o.SpatialScene()
    fr = [o.Frames.from_slice(48000, p) for p in pcms]
    for s in range(n_src):
        full.play(o.FramesSignal(fr[s % 4], 1.0), pos[s], vel[s], 0.1)
    for k in range(n_cb):
        ref = o.run(full, 48000, n_frames)
        for r in range(1, world):
            assert np.array_equal(got[0][k], got[r][k]), "ranks must hold bit-identical sums"
        assert_mix_close(got[0][k], ref, ref.astype(np.float64))


def test_single_rank_in_kernel_exchange_equals_plain_sample(oracle):
    """World size 1: odb_scene_sample_exchange is the plain callback plus the epilogue on the (one-term) sum."""
    import torch

    import oddio_b200 as odb
    from helpers import rand_in_shell, synth_pcm
    from oddio_b200.sharding import PeerExchange

    ctx = odb.init(0)
    rng = np.random.default_rng(5)
    pcms = [synth_pcm(rng, 60000, 48000) for _ in range(3)]
    frames = [odb.Frames.from_slice(48000, p, ctx) for p in pcms]
    pos = [rand_in_shell(rng, 2, 60) for _ in range(50)]
    vel = [rng.uniform(-30, 30, 3).astype(np.float32) for _ in range(50)]

    def scene():
        ctl, sc = odb.SpatialScene.new(ctx)
        for i in range(50):
            ctl.play(odb.FramesSignal(frames[i % 3], 1.0), odb.SpatialOptions(pos[i], vel[i], 0.1))
        return ctl, sc

    interval = float(np.float32(1.0) / np.float32(48000))
    for n_frames, epi in ((1024, 0), (700, 1), (2048, 2)):
        ex = PeerExchange(ctx, 0, 1, 2 * n_frames, depth=2)
        (_, a), (_, b) = scene(), scene()
        tile = torch.zeros((n_frames, 2), device="cuda", dtype=torch.float32)
        for _ in range(3):
            want = a.sample(interval, n_frames).astype(np.float64)
            assert b.sample_exchange(ex, interval, tile.data_ptr(), n_frames, lag=0, epilogue=epi)
            ctx.synchronize()
            got = tile.cpu().numpy()
            if epi == 0:
                assert np.array_equal(got, want.astype(np.float32))
            elif epi == 1:
                np.testing.assert_allclose(got, np.tanh(want), rtol=0, atol=4e-7)
            else:
                assert np.array_equal(got, (want.astype(np.float32) / (np.float32(1.0) + np.abs(want.astype(np.float32)))))
        a.close(); b.close(); ex.close()


def test_in_kernel_exchange_entry_point_with_buffered_sources_falls_back_to_the_stand_alone_kernels(oracle):
    """A scene with play_buffered sources takes the multi-kernel callback; odb_scene_sample_exchange then exchanges
    with the stand-alone push / pull kernels - same result as the plain callback (world size 1), same lag protocol."""
    import torch

    import oddio_b200 as odb
    from helpers import rand_in_shell, synth_pcm
    from oddio_b200.sharding import PeerExchange

    ctx = odb.init(0)
    rng = np.random.default_rng(9)
    pcm = synth_pcm(rng, 90000, 48000)
    fr = odb.Frames.from_slice(48000, pcm, ctx)
    pos = [rand_in_shell(rng, 2, 60) for _ in range(12)]
    vel = [rng.uniform(-20, 20, 3).astype(np.float32) for _ in range(12)]

    def scene():
        ctl, sc = odb.SpatialScene.new(ctx)
        for i in range(12):
            if i % 3 == 0:
                ctl.play_buffered(odb.FramesSignal(fr, 0.0), odb.SpatialOptions(pos[i], vel[i], 0.1), 100.0, 48000, 0.1)
            else:
                ctl.play(odb.FramesSignal(fr, 1.0), odb.SpatialOptions(pos[i], vel[i], 0.1))
        return ctl, sc

    interval, n = float(np.float32(1.0) / np.float32(48000)), 1024
    ex = PeerExchange(ctx, 0, 1, 2 * n, depth=3)
    (_, a), (_, b) = scene(), scene()
    tile = torch.zeros((n, 2), device="cuda", dtype=torch.float32)
    want = [a.sample(interval, n).copy() for _ in range(4)]
    got = []
    for k in range(4):
        if b.sample_exchange(ex, interval, tile.data_ptr(), n, lag=1):
            ctx.synchronize()
            got.append(tile.cpu().numpy().copy())
    ex.pull(tile.data_ptr(), 2 * n)
    ctx.synchronize()
    got.append(tile.cpu().numpy().copy())
    assert len(got) == 4
    for k in range(4):
        assert np.array_equal(got[k], want[k])
    assert np.abs(want[-1]).max() > 0
    a.close(); b.close(); ex.close()
