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


def _setup_failure_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oddio_b200 as odb
    from oddio_b200.sharding import PeerExchange

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class NoContext:  # stands in for a rank whose CUDA context could not be created
        _h = None

    try:
        PeerExchange.from_torch(NoContext(), 2048, depth=4)
        q.put((rank, "returned"))
    except odb.OddioError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_setup_failure_is_agreed_by_all_ranks():
    """PeerExchange.from_torch without a usable device (this box): every rank still takes part in both set-up
    collectives, nobody hangs, and every rank gets the same OddioError naming the ranks that failed - which is
    what lets bench.py fall back to the NCCL exchange on all ranks together."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_setup_failure_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(got) == {0, 1}
    for msg in got.values():
        assert "peer-memory exchange set-up failed" in msg and "rank 0" in msg and "rank 1" in msg


class _FakeExchangeLib:
    """The odb_exchange_* entry points as a host-side stand-in, to drive PeerExchange's set-up protocol on a box
    without a GPU: handles are 64 bytes carrying the exporting rank, connect records what it was given."""

    def __init__(self, log):
        self.log = log

    def odb_exchange_create(self, ctx_h, rank, world, max_floats, depth, out):
        self.rank, self.world = rank, world
        out._obj.value = 0x1000 + rank
        self.log.append(("create", rank, world, max_floats, depth))
        return 0

    def odb_exchange_handle_size(self):
        return 64

    def odb_exchange_export(self, h, buf):
        buf.raw = bytes([self.rank + 1]) * 64
        return 0

    def odb_exchange_connect(self, h, blob):
        self.log.append(("connect", [blob.raw[64 * r] for r in range(self.world)]))
        return 0

    def odb_exchange_destroy(self, h):
        self.log.append(("destroy",))
        return 0


def _setup_success_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oddio_b200 import _lib
    from oddio_b200.sharding import PeerExchange

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    log = []
    fake = _FakeExchangeLib(log)
    _lib.load = lambda: fake

    class Ctx:
        _h = None

    ex = PeerExchange.from_torch(Ctx(), 16384, depth=4)
    ex.close()
    q.put((rank, log))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_setup_protocol_over_torch_distributed():
    """The success path of PeerExchange.from_torch with the C entry points replaced by a recorder: every rank creates
    its inbox with the requested depth, exports once, and connects with the handles of all ranks in rank order."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 3
    procs = [ctx.Process(target=_setup_success_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r] == [("create", r, world, 16384, 4), ("connect", [1, 2, 3]), ("destroy",)]
