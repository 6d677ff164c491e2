"""Multi-GPU plumbing: one process per GPU, particles sharded by index range, grid moments
all-reduced inside libgempic_b200 (NCCL over NVLink) -- the B200 counterpart of the
reference's `Threads.@spawn` chunks + `reduce(+, fetch.(tasks))`
(src/hamiltonian_splitting.jl:61-66, src/hamiltonian_splitting_1d2v.jl:48-88).

torch.distributed is used for the rendezvous only (broadcast of the NCCL unique id, barriers,
max-over-ranks timing); the data path never goes through Python.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def shard_range(n_global: int, world_size: int, rank: int):
    """Contiguous index range [first, first+count) of `rank`: ceil(N/G) particles per rank, the
    last ranks take the remainder (the reference's `@assert np % nthreads() == 0` is dropped)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0,{world_size})")
    per = -(-n_global // world_size)
    first = min(rank * per, n_global)
    count = min(per, n_global - first)
    return first, count


class DistributedContext:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and owns the process group."""

    def __init__(self, backend: str | None = None, init_process_group: bool = True):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.backend = backend
        self._dist = None
        self._comm_ready = False
        if self.world_size > 1 and init_process_group:
            import torch
            import torch.distributed as dist

            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            self.backend = backend
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)
            self._dist = dist

    # -- helpers that work for any backend -----------------------------------------------------
    def _tensor(self, arr):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.cuda(self.local_rank) if self.backend == "nccl" else t

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()

    def broadcast_bytes(self, payload: bytes | None, n: int, src: int = 0) -> bytes:
        """Broadcast an n-byte blob from `src` to every rank."""
        if self._dist is None:
            return payload
        import torch

        buf = np.frombuffer(payload, dtype=np.uint8).copy() if self.rank == src else np.zeros(n, dtype=np.uint8)
        t = self._tensor(buf)
        self._dist.broadcast(t, src=src)
        return bytes(t.cpu().numpy().tobytes())

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        if self._dist is None:
            return np.array(arr, copy=True)
        t = self._tensor(np.asarray(arr, dtype=np.float64))
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def max_over_ranks(self, value: float) -> float:
        if self._dist is None:
            return float(value)
        t = self._tensor(np.array([value], dtype=np.float64))
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.cpu().numpy()[0])

    def shard(self, n_global: int):
        return shard_range(n_global, self.world_size, self.rank)

    # -- the library communicator -----------------------------------------------------------------
    def init_library_comm(self):
        """Create the NCCL communicator inside libgempic_b200 (rank 0 makes the id)."""
        from . import _lib

        _lib.init(self.local_rank)
        if self.world_size == 1 or self._comm_ready:
            return
        L = _lib.load()
        blob = None
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(L.gempic_comm_unique_id(buf))
            blob = bytes(buf)
        blob = self.broadcast_bytes(blob, 128, src=0)
        _lib.check(L.gempic_comm_init(C.c_int(self.world_size), C.c_int(self.rank), C.c_char_p(blob)))
        self._comm_ready = True

    def finalize(self):
        from . import _lib

        if self._comm_ready:
            _lib.check(_lib.load().gempic_comm_finalize())
            self._comm_ready = False
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
            self._dist = None
