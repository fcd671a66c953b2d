"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` only for plumbing.

* Batched 1-D transforms, Clenshaw point sets and sampling shard by contiguous ranges with NO collective
  (`shard_range`): SURVEY.md section 8(e).
* The 2-D tensor transform has one real exchange: column slabs -> all-to-all transpose -> row slabs
  (`Sharded2DPlan`, NCCL grouped send/recv inside libapproxfun_b200.so).  The index arithmetic of the
  pack / exchange / unpack is mirrored by `slab_exchange_reference` so it can be tested with the `gloo`
  backend on CPUs (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) owned by `rank`; earlier ranks get the remainder (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def slab_exchange_reference(local_cols: np.ndarray, rank: int, world: int, send, recv) -> np.ndarray:
    """NumPy model of the all-to-all transpose used by afb_sharded2d_execute_device.

    local_cols: this rank's column block AFTER the column transforms, shape (n, n2/world) (element [i, j]).
    send(dst, array) / recv(src, shape) move contiguous float64 blocks between ranks.
    Returns this rank's row block as (n/world, n2): row i_local holds all n2 columns.
    """
    n, nc = local_cols.shape
    nr = n // world
    # pack = local transpose: T[i*nc + j]; rows of destination d are the contiguous range [d*nr, (d+1)*nr)
    T = np.ascontiguousarray(local_cols)  # C order (n, nc) is exactly T[i*nc + j]
    reqs = []
    for d in range(world):
        if d != rank:
            reqs.append(send(d, T[d * nr:(d + 1) * nr]))
    R = np.empty((world, nr, nc))
    R[rank] = T[rank * nr:(rank + 1) * nr]
    for s in range(world):
        if s != rank:
            R[s] = recv(s, (nr, nc))
    for r in reqs:
        if r is not None:
            r.wait()
    # unpack: L[i][s*nc + jj] = R[s][i][jj]
    return np.ascontiguousarray(np.transpose(R, (1, 0, 2)).reshape(nr, world * nc))


class Sharded2DPlan:
    """2-D Chebyshev transform of an n x n2 matrix sharded by column blocks over the process group.

    in : device pointer to this rank's block, column-major n x (n2/world)
    out: device pointer to this rank's n/world rows of the result, each a contiguous line of length n2
    """

    _comm_ready = False

    def __init__(self, kind: int, n: int, n2: int, p2p: bool = True):
        import torch
        import torch.distributed as dist

        from .api import lowlevel

        self.lib = lowlevel().lib
        rank, world = dist.get_rank(), dist.get_world_size()
        if not Sharded2DPlan._comm_ready:
            ident = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_ubyte * 128)()
                L.check(self.lib, self.lib.afb_comm_unique_id(buf))
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            ident = ident.to(dev)
            dist.broadcast(ident, 0)
            raw = bytes(ident.cpu().tolist())
            L.check(self.lib, self.lib.afb_comm_create(raw, rank, world))
            Sharded2DPlan._comm_ready = True
        h = C.c_void_p()
        L.check(self.lib, self.lib.afb_sharded2d_create(C.byref(h), kind, n, n2))
        self._h = h
        self.n, self.n2, self.world, self.rank = n, n2, world, rank
        self.p2p = False
        if p2p and world > 1 and dist.get_backend() == "nccl":
            # peer-memory mode: all-gather the 64-byte CUDA IPC handles of the receive buffers
            mine = (C.c_ubyte * 64)()
            L.check(self.lib, self.lib.afb_sharded2d_ipc_export(self._h, mine))
            dev = torch.device("cuda", torch.cuda.current_device())
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            allh = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allh, t)
            blob = b"".join(bytes(x.cpu().tolist()) for x in allh)
            L.check(self.lib, self.lib.afb_sharded2d_ipc_open(self._h, blob, world))
            self.p2p = True

    def set_chunks(self, chunks: int):
        """Peer-memory mode: column pass in `chunks` pieces, each sent on a side stream while the next is transformed."""
        L.check(self.lib, self.lib.afb_sharded2d_set_chunks(self._h, int(chunks)))

    def set_tuning(self, order: int = 0, grid_mult: int = 8, novec: int = 0):
        L.check(self.lib, self.lib.afb_sharded2d_set_tuning(self._h, int(order), int(grid_mult), int(novec)))

    def set_phases(self, mask: int):
        """Measurement only: bit 0 column pass, bit 1 exchange + wait, bit 2 row pass (7 = everything)."""
        L.check(self.lib, self.lib.afb_sharded2d_set_phases(self._h, int(mask)))

    def execute(self, in_ptr: int, out_ptr: int, stream: int = 0):
        L.check(self.lib, self.lib.afb_sharded2d_execute_device(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), C.c_void_p(stream)))

    def close(self):
        if self._h is not None:
            L.check(self.lib, self.lib.afb_sharded2d_destroy(self._h))
            self._h = None
