"""Sweeps the exchange-kernel variants of the sharded 2-D transform (torchrun, one rank per GPU; not a benchmark of record).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/time_sharded2d.py [n]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402
from approxfun_b200 import parallel as par  # noqa: E402

L.init(L.load(), local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
st = torch.cuda.current_stream().cuda_stream
nc = n // world
blk = torch.randn(nc, n, dtype=torch.float64, device="cuda")
rows = torch.empty(n // world, n, dtype=torch.float64, device="cuda")


def barrier():
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


sp = par.Sharded2DPlan(L.CHEB1_2D_FWD, n, n, p2p=True)
nv = 8.0 * n * nc * (world - 1) / world
for order in (3, 4):
    for gm in (1, 2, 3, 8):
        for novec in (0, 1):
            if novec:
                continue
            sp.set_tuning(order, gm, novec)
            sp.set_phases(2)
            tx = timeit(lambda: sp.execute(blk.data_ptr(), rows.data_ptr(), st))
            sp.set_phases(7)
            tt = timeit(lambda: sp.execute(blk.data_ptr(), rows.data_ptr(), st))
            if rank == 0:
                print(f"order {order} grid {gm}x novec {novec}: exchange {tx:.3f} ms ({nv / tx / 1e6:.0f} GB/s)  total {tt:.3f} ms", flush=True)
sp.set_tuning(3, 8, 0)
for ch in (2, 4):
    sp.set_chunks(ch)
    tt = timeit(lambda: sp.execute(blk.data_ptr(), rows.data_ptr(), st))
    if rank == 0:
        print(f"chunks {ch}: total {tt:.3f} ms", flush=True)
barrier()
sp.close()
dist.destroy_process_group()
