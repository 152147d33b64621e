"""Multi-GPU plumbing of the spatial/mixer path: one process per GPU, sources sharded over the ranks,
one sum of the small output tile per callback. The reference has no counterpart (it is
single-process). On GPUs the exchange is `PeerExchange`: the library's own one-kernel push/sum over NVLink
peer memory (csrc/odb_exchange.cu); torch.distributed only carries the 64-byte set-up handles (and the
NCCL / gloo all-reduce of `allreduce_tile`, kept for comparison and for the CPU tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_sources(n_sources: int, rank: int, world: int) -> np.ndarray:
    """Round-robin: source s belongs to rank s mod world (keeps the ranks balanced as sources finish)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_sources, world, dtype=np.int64)


def allreduce_tile(tile, group=None):
    """Sums the per-rank partial tiles in place; `tile` is a torch tensor on the rank's device."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tile, op=dist.ReduceOp.SUM, group=group)
    return tile


class PeerExchange:
    """Sum of the per-rank tiles over NVLink peer memory (include/oddio_b200.h, odb_exchange_*).

    `gather(handle: bytes) -> list[bytes]` returns every rank's handle in rank order; `from_torch` builds
    it from torch.distributed (any backend: the handles are host bytes). `depth`: pushed exchanges that may
    await their pull (2..8)."""

    def __init__(self, ctx, rank: int, world: int, max_floats: int, gather=None, depth: int = 2):
        from . import _lib

        self._lib = _lib
        h = C.c_void_p()
        _lib.check(_lib.load().odb_exchange_create(ctx._h, int(rank), int(world), int(max_floats), int(depth), C.byref(h)))
        self._h = h
        self.rank, self.world = rank, world
        if world > 1:
            if gather is None:
                raise ValueError("world > 1 needs a gather function for the set-up handles")
            if gather is not False:  # False: the caller drives _export / _connect itself (from_torch)
                self._connect(gather(self._export()))

    def _export(self) -> bytes:
        L = self._lib.load()
        buf = C.create_string_buffer(L.odb_exchange_handle_size())
        self._lib.check(L.odb_exchange_export(self._h, buf))
        return buf.raw

    def _connect(self, handles) -> None:
        L = self._lib.load()
        n = L.odb_exchange_handle_size()
        if len(handles) != self.world or any(len(x) != n for x in handles):
            raise ValueError("gather must return one handle per rank, in rank order")
        blob = C.create_string_buffer(b"".join(handles), n * self.world)
        self._lib.check(L.odb_exchange_connect(self._h, blob))

    @classmethod
    def from_torch(cls, ctx, max_floats: int, group=None, depth: int = 2):
        """Builds the exchange over the ranks of a torch.distributed group. Every rank takes part in both collectives
        below whether or not its own set-up succeeded, and all ranks agree on the outcome: the connected exchange is
        returned on every rank, or OddioError is raised on every rank (callers can then fall back together)."""
        import torch.distributed as dist

        from . import _lib

        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return cls(ctx, 0, 1, max_floats, depth=depth)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ex, err, handle = None, None, b""
        try:
            ex = cls(ctx, rank, world, max_floats, gather=False, depth=depth)  # local part: inbox + handle
            handle = ex._export()
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            err = f"rank {rank}: {e}"
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        if err is None:
            try:
                if not all(handles):
                    raise RuntimeError("a peer has no inbox")
                ex._connect(handles)
            except Exception as e:  # noqa: BLE001
                err = f"rank {rank}: {e}"
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)  # doubles as the barrier: every rank has mapped every inbox
        bad = [e for e in errs if e]
        if bad:
            if ex is not None:
                ex.close()
            raise _lib.OddioError(_lib.ODB_E_CUDA, "peer-memory exchange set-up failed: " + "; ".join(bad))
        return ex

    def allreduce(self, dev_ptr: int, n_floats: int, epilogue: int = 0, stream: int = 0) -> None:
        """In place on the device tile at `dev_ptr`; queued on `stream` (0 = the context's stream)."""
        self._lib.check(self._lib.load().odb_exchange_allreduce(self._h, C.c_void_p(dev_ptr), int(n_floats), int(epilogue),
                                                                C.c_void_p(stream) if stream else None))

    def push(self, dev_ptr: int, n_floats: int, stream: int = 0) -> None:
        """First half of allreduce: sends this rank's tile to every rank; never waits for a peer's data."""
        self._lib.check(self._lib.load().odb_exchange_push(self._h, C.c_void_p(dev_ptr), int(n_floats),
                                                           C.c_void_p(stream) if stream else None))

    def pull(self, dev_ptr: int, n_floats: int, epilogue: int = 0, stream: int = 0) -> None:
        """Second half, in push order (at most `depth` pushes outstanding): the sum over the ranks lands in `dev_ptr`."""
        self._lib.check(self._lib.load().odb_exchange_pull(self._h, C.c_void_p(dev_ptr), int(n_floats), int(epilogue),
                                                           C.c_void_p(stream) if stream else None))

    def close(self) -> None:
        if self._h:
            self._lib.load().odb_exchange_destroy(self._h)
            self._h = None
