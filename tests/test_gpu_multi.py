"""Multi-GPU parity (needs >= 2 B200s): sharding by batch must be bit-identical to the 1-GPU result, and the
slab-decomposed 2-D transform (NCCL all-to-all over NVLink) must equal the 1-GPU 2-D transform."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
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
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import approxfun_b200 as af
        from approxfun_b200 import _lib as L
        from approxfun_b200 import parallel as par

        lib = L.load()
        L.init(lib, rank)
        st = torch.cuda.current_stream().cuda_stream
        # (1) batch sharding, no collective on the data path
        n, B = 4096, 1024
        g = torch.Generator(device="cuda").manual_seed(7)
        V = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g)   # same on every rank
        whole = torch.empty_like(V)
        af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n).execute(V.data_ptr(), whole.data_ptr(), st)
        lo, hi = par.shard_range(B, rank, world)
        mine = torch.empty(hi - lo, n, dtype=torch.float64, device="cuda")
        af.api.DevicePlan(L.CHEB1_FWD, n, 0, hi - lo, n).execute(V[lo:hi].data_ptr(), mine.data_ptr(), st)
        torch.cuda.synchronize()
        ok1 = bool(torch.equal(mine, whole[lo:hi]))
        # (2) 2-D slab transform with the all-to-all
        n1, n2 = 1024, 2048
        M = torch.randn(n2, n1, dtype=torch.float64, device="cuda", generator=g)   # column-major n1 x n2
        full = torch.empty_like(M)
        af.api.DevicePlan(L.CHEB1_2D_FWD, n1, n2, 1, n1).execute(M.data_ptr(), full.data_ptr(), st)
        nc, nr = n2 // world, n1 // world
        block = M[rank * nc:(rank + 1) * nc].contiguous()
        # full is column-major n1 x n2: full[j, i] = C[i, j]; out[i_local, j] = C[rank*nr + i_local, j]
        want = full[:, rank * nr:(rank + 1) * nr].t().contiguous()
        errs = []
        for p2p in (False, True):     # NCCL send/recv all-to-all, then the fused peer-memory transpose kernel
            out = torch.zeros(nr, n2, dtype=torch.float64, device="cuda")
            plan = par.Sharded2DPlan(L.CHEB1_2D_FWD, n1, n2, p2p=p2p)
            assert plan.p2p == p2p
            for _ in range(3):        # repeated calls exercise the write-after-read barrier
                plan.execute(block.data_ptr(), out.data_ptr(), st)
            torch.cuda.synchronize()
            errs.append(float((out - want).abs().max()))
            if p2p:
                # every exchange-kernel variant: tile orders, the TMA bulk-store kernel (order 4), chunked column pass
                for order, chunks in ((0, 1), (1, 1), (2, 1), (4, 1), (3, 2), (4, 4)):
                    plan.set_tuning(order, 3, 0)
                    plan.set_chunks(chunks)
                    out.zero_()
                    for _ in range(2):
                        plan.execute(block.data_ptr(), out.data_ptr(), st)
                    torch.cuda.synchronize()
                    errs.append(float((out - want).abs().max()))
            dist.barrier()
            plan.close()
        q.put((rank, ok1, max(errs)))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_and_alltoall():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res          # bit-identical under batch sharding
    assert all(r[2] < 1e-13 for r in res), res  # 2-D all-to-all path equals the 1-GPU path


def test_single_process_two_gpus_through_the_c_abi():
    """afb_init(devices, 2): ONE process (what a Julia session is) drives two GPUs; the host-pointer calls shard inside
    the library (host threads + per-device streams, peer-memory transpose for the 2-D plan; no torch.distributed, no IPC)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import approxfun_b200 as af
    from approxfun_b200 import _lib as L
    from oracle import approxfun_oracle as orc
    from oracle import oracle_c as orcc

    rng = np.random.default_rng(3)
    n, B = 4096, 4096
    v = rng.standard_normal((B, n))
    M = rng.standard_normal((2048, 1024))            # column-major 1024 x 2048
    c10k = rng.standard_normal(3000) / (1 + np.arange(3000))
    x = rng.uniform(-1, 1, 1 << 21)
    f = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
    cdf = af.normalizedcumsum(f).coefficients
    u = rng.random(1 << 22)
    plan = af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n)
    plan2 = af.api.DevicePlan(L.CHEB1_2D_FWD, 1024, 2048, 1, 1024)
    ll = af.api.lowlevel()
    try:
        af.api.use_devices(0)
        one, one2 = np.empty_like(v), np.empty_like(M)
        plan.execute_host(v, one)
        plan2.execute_host(M, one2)
        y1 = ll.clenshaw(c10k, x)
        s1 = ll.samplecdf(cdf, -5.0, 5.0, u)
        r1 = ll.sample_cheb(cdf, -5.0, 5.0, 99, 1 << 22)
        af.api.use_devices([0, 1])
        two, two2 = np.empty_like(v), np.empty_like(M)
        for _ in range(3):                           # repeated calls alternate the 2-D receive buffers
            plan.execute_host(v, two)
            plan2.execute_host(M, two2)
            assert np.array_equal(one, two)          # batch sharding: bit-identical to one device
            assert np.max(np.abs(one2 - two2)) <= 1e-13 * 11 * np.max(np.abs(M))
        assert np.max(np.abs(two2.T - orc.cheb2_transform(M.T))) <= 1e-13 * 11 * np.max(np.abs(M))
        assert np.array_equal(ll.clenshaw(c10k, x), y1)
        assert np.array_equal(ll.samplecdf(cdf, -5.0, 5.0, u), s1)
        assert np.array_equal(ll.sample_cheb(cdf, -5.0, 5.0, 99, 1 << 22), r1)   # Philox counters do not depend on the device count
        assert np.array_equal(y1[:5000], orcc.clenshaw(c10k, x[:5000]))
    finally:
        af.api.use_devices(0)
        plan.close()
        plan2.close()
