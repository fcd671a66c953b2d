"""Cost of cudaHostRegister / cudaHostUnregister on a pageable 2 GiB array, whole and in chunks, 1 and 2 threads
(behind the choice between the pinned ring and transient registration in afb_plan_execute_host; not a benchmark of record)."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402

lib = af.api.lowlevel().lib
a = np.ones(1 << 28)          # 2 GiB, touched
b = np.ones(1 << 28)
base_a, base_b = a.ctypes.data, b.ctypes.data


def reg(ptr, nbytes):
    st = lib.afb_host_register(C.c_void_p(ptr), C.c_size_t(nbytes))
    assert st == 0, st


def unreg(ptr):
    st = lib.afb_host_unregister(C.c_void_p(ptr))
    assert st == 0, st


for chunk_mb in (2048, 256, 64, 32):
    ch = chunk_mb << 20
    nchunk = a.nbytes // ch
    t0 = time.perf_counter()
    for i in range(nchunk):
        reg(base_a + i * ch, ch)
    t1 = time.perf_counter()
    for i in range(nchunk):
        unreg(base_a + i * ch)
    t2 = time.perf_counter()
    print(f"1 thread, chunks of {chunk_mb} MiB: register {1e3 * (t1 - t0):.1f} ms, unregister {1e3 * (t2 - t1):.1f} ms per 2 GiB", flush=True)

for chunk_mb in (64,):
    ch = chunk_mb << 20
    nchunk = a.nbytes // ch

    def work(base):
        for i in range(nchunk):
            reg(base + i * ch, ch)
        for i in range(nchunk):
            unreg(base + i * ch)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(x,)) for x in (base_a, base_b)]
    [t.start() for t in th]
    [t.join() for t in th]
    print(f"2 threads (one array each), chunks of {chunk_mb} MiB: register + unregister of 2 x 2 GiB in {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
