"""Sharded (multi-GPU) front end: one process per GPU, `torch.distributed` only for the
plumbing (rendezvous, exchanging the CUDA-IPC handle blobs, host barriers, max-over-ranks
timing).  The data path has no collective call: the per-pivot exchanges are peer stores
issued by the persistent kernel itself (csrc/kernels.cuh, "sharded loop").

Partition: B^-1 by row blocks, A_N by column blocks, O(m) vectors replicated.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .solver import Engine, _dtype_code, _vp


@dataclass(frozen=True)
class ShardPlan:
    """Host mirror of the engine's partition arithmetic (Engine ctor / set_columns in engine.cu)."""
    m: int
    n: int
    world: int
    itemsize: int = 8

    @property
    def ld(self) -> int:
        q = 32 * (16 // self.itemsize)            # one warp-wide 16-byte vector row
        return (self.m + q - 1) // q * q

    def rows(self, rank: int):
        """[row0, row1) of the padded B^-1 owned by `rank` (multiples of 64 doubles / 128 floats)."""
        q = 32 * (16 // self.itemsize)
        rpr = ((self.ld + self.world - 1) // self.world + q - 1) // q * q
        return min(self.ld, rank * rpr), min(self.ld, (rank + 1) * rpr)

    def cols(self, rank: int):
        """[col0, col1) of the n - m structural columns owned by `rank`."""
        ns = self.n - self.m
        return ns * rank // self.world, ns * (rank + 1) // self.world

    def slack_cols(self, rank: int):
        return self.m * rank // self.world, self.m * (rank + 1) // self.world

    def owner_of_col(self, p: int) -> int:
        for r in range(self.world):
            c0, c1 = self.cols(r)
            if c0 <= p < c1:
                return r
        raise ValueError("slack columns have no owner (unit vectors)")

    def owner_of_row(self, q: int) -> int:
        for r in range(self.world):
            r0, r1 = self.rows(r)
            if r0 <= q < r1:
                return r
        raise ValueError(q)

    def bytes_per_pivot(self, rank: int) -> int:
        """Algorithmic HBM bytes of one pivot on `rank`: its B^-1 rows read + written, its A columns read."""
        r0, r1 = self.rows(rank)
        c0, c1 = self.cols(rank)
        rows = max(0, min(self.m, r1) - r0)
        return self.itemsize * (2 * rows * self.m + self.m * (c1 - c0))

    def exchange_bytes_per_pivot(self) -> int:
        """NVLink payload one rank SENDS per pivot: candidate + its alpha slice + (owner only) row q, to every peer."""
        rows = self.rows(0)[1] - self.rows(0)[0]
        return (self.world - 1) * (16 + self.itemsize * rows) + (self.world - 1) * self.itemsize * self.m // self.world


def exchange_blobs(blob: bytes, group=None, device=None) -> bytes:
    """All-gather one fixed-size byte blob per rank, in rank order (works on gloo/CPU and nccl/CUDA)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        src = src.to(device)
    out = torch.empty(world * len(blob), dtype=torch.uint8, device=src.device)
    dist.all_gather_into_tensor(out, src, group=group) if src.is_cuda else \
        dist.all_gather(list(out.view(world, len(blob)).unbind(0)), src, group=group)
    return bytes(out.cpu().numpy().tobytes())


class ShardedEngine(Engine):
    """One rank of the sharded engine.  Every rank must make the same calls in the same order."""

    def __init__(self, m: int, n: int, dtype=np.float64, rank: int | None = None, world: int | None = None,
                 group=None, device: int | None = None, **opts):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.m, self.n, self.dtype = int(m), int(n), np.dtype(dtype)
        self.plan = ShardPlan(self.m, self.n, self.world, self.dtype.itemsize)
        self._L = capi.lib()
        dev = self.rank if device is None else device
        self._o = capi.default_options(device=dev, **opts)
        self._h = C.c_void_p()
        capi.check(self._L.b200lp_create_sharded(_dtype_code(self.dtype), self.m, self.n, self.rank, self.world,
                                                 C.byref(self._o), C.byref(self._h)))
        self.device = dev

    def connect(self):
        """Exchange the IPC handles of the A shards / mailboxes and map the peers (after upload/generate)."""
        import torch
        nb = self._L.b200lp_ipc_handle_bytes()
        mine = C.create_string_buffer(nb)
        capi.check(self._L.b200lp_ipc_export(self._h, mine))
        allb = exchange_blobs(mine.raw, self.group, torch.device("cuda", self.device))
        capi.check(self._L.b200lp_ipc_import(self._h, allb, self.world))

    def upload_columns(self, A_cols, b, c):
        """A_cols: this rank's structural columns only (m x ncols, column-major)."""
        c0, c1 = self.plan.cols(self.rank)
        A_cols = np.asfortranarray(A_cols, dtype=self.dtype)
        assert A_cols.shape == (self.m, c1 - c0)
        b = np.ascontiguousarray(b, dtype=self.dtype)
        c = np.ascontiguousarray(c, dtype=self.dtype)
        capi.check(self._L.b200lp_upload_columns(self._h, _vp(A_cols), c0, c1 - c0, _vp(b), _vp(c)))

    def shard_rows(self):
        r0, rows = C.c_int64(0), C.c_int64(0)
        capi.check(self._L.b200lp_shard_rows(self._h, C.byref(r0), C.byref(rows)))
        return r0.value, rows.value

    def download_binv_rows(self) -> np.ndarray:
        """This rank's rows of B^-1 (rows x m)."""
        _, rows = self.shard_rows()
        B = np.zeros((rows, self.m), self.dtype, order="F")
        capi.check(self._L.b200lp_download_binv(self._h, _vp(B)))
        return B
