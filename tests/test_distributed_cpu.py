"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: batch sharding and the slab all-to-all transpose."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import approxfun_b200 as af
        from approxfun_b200 import parallel as par
        from oracle import approxfun_oracle as orc

        # (1) batch sharding without a collective: every rank transforms its slice, results tile the whole
        n, B = 64, 37
        rng = np.random.default_rng(0)
        V = rng.standard_normal((n, B))
        lo, hi = par.shard_range(B, rank, world)
        mine = orc.cheb_transform(V[:, lo:hi], axis=0)          # stands in for the GPU kernel on this rank
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, mine))
        whole = np.concatenate([p[2] for p in sorted(parts, key=lambda t: t[0])], axis=1)
        ok1 = np.array_equal(whole, orc.cheb_transform(V, axis=0)) or np.max(np.abs(whole - orc.cheb_transform(V, axis=0))) < 1e-15
        cover = sorted((p[0], p[1]) for p in parts)
        ok1 = ok1 and cover[0][0] == 0 and cover[-1][1] == B and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))

        # (2) 2-D: column slabs -> column transforms -> all-to-all transpose -> row transforms
        n, n2 = 16, 24
        V2 = rng.standard_normal((n, n2))
        nc = n2 // world
        local = orc.cheb_transform(V2[:, rank * nc:(rank + 1) * nc], axis=0)

        def send(dst, arr):
            return dist.isend(torch.from_numpy(np.ascontiguousarray(arr)), dst)

        def recv(src, shape):
            t = torch.empty(shape, dtype=torch.float64)
            dist.recv(t, src)
            return t.numpy()

        rows = par.slab_exchange_reference(local, rank, world, send, recv)    # (n/world, n2)
        rows = orc.cheb_transform(rows, axis=1)
        want = orc.cheb2_transform(V2)[rank * (n // world):(rank + 1) * (n // world), :]
        ok2 = np.max(np.abs(rows - want)) < 1e-14
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_sharding_and_slab_transpose():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_shard_range_properties():
    import sys

    sys.path.insert(0, ROOT)
    from approxfun_b200.parallel import shard_range

    for total in (0, 1, 7, 65536, 10 ** 9):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)
