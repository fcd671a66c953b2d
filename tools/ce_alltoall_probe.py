"""Copy-engine all-to-all probe (one process, all GPUs of the node): every GPU sends its share of a 16384 x (16384/G) column
slab to every peer with cudaMemcpy2DAsync (pitched, `width` bytes per row) on one stream per destination -- the transfer an
SM-free, transpose-free exchange of the sharded 2-D transform would make.  Prints the per-GPU outgoing rate.
Not a benchmark of record.   usage: python tools/ce_alltoall_probe.py [n]"""
import sys
import time

import torch
from cuda.bindings import runtime as cudart

G = torch.cuda.device_count()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cols = n // G                     # columns owned per GPU
rows = n // G                     # rows of each destination's block


def chk(r):
    err = r[0] if isinstance(r, tuple) else r
    assert int(err) == 0, err
    return r[1] if isinstance(r, tuple) and len(r) > 1 else None


for d in range(G):
    torch.cuda.set_device(d)
    for p in range(G):
        if p != d:
            r = cudart.cudaDeviceEnablePeerAccess(p, 0)
            assert int(r[0]) in (0, 704), r      # 704: already enabled

src = [torch.randn(cols, n, dtype=torch.float64, device=f"cuda:{d}") for d in range(G)]          # cols columns of n doubles
dst = [torch.empty(n, rows, dtype=torch.float64, device=f"cuda:{d}") for d in range(G)]          # all n columns x my rows
streams = [[torch.cuda.Stream(device=d) for _ in range(G)] for d in range(G)]
D2D = cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice


def alltoall(nbands, colchunks):
    """bands of rows/nbands rows (width = 8 * rows / nbands bytes), column chunks of cols/colchunks columns"""
    bw = rows // nbands
    cc = cols // colchunks
    for k in range(colchunks):
        for d in range(G):
            torch.cuda.set_device(d)
            for s in range(1, G):
                p = (d + s) % G
                st = streams[d][p].cuda_stream
                for b in range(nbands):
                    # destination layout [band][column][bw rows]: band b of peer p, columns d*cols + k*cc ...
                    dptr = dst[p].data_ptr() + 8 * (b * n * bw + (d * cols + k * cc) * bw)
                    sptr = src[d].data_ptr() + 8 * (k * cc * n + p * rows + b * bw)
                    chk(cudart.cudaMemcpy2DAsync(dptr, 8 * bw, sptr, 8 * n, 8 * bw, cc, D2D, st))


def sync():
    for d in range(G):
        torch.cuda.synchronize(d)


print(f"{G} GPUs, n = {n}: {8 * cols * (n - rows) / 1e6:.0f} MB out per GPU")
for nbands, colchunks in ((1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (4, 1)):
    alltoall(nbands, colchunks)
    sync()
    best = 1e9
    REP = 8                       # queue REP sets back to back so that the host's issue time hides behind the copies
    for _ in range(3):
        sync()
        t0 = time.perf_counter()
        for _r in range(REP):
            alltoall(nbands, colchunks)
        t_issue = time.perf_counter() - t0
        sync()
        best = min(best, (time.perf_counter() - t0) / REP)
    print(f"   (host issue time per set: {t_issue / REP * 1e3:.3f} ms)")
    print(f"bands {nbands} (rows of {8 * rows // nbands} B), column chunks {colchunks}: {best * 1e3:.3f} ms wall "
          f"= {8 * cols * (n - rows) / best / 1e9:.0f} GB/s per GPU each way", flush=True)
# check one block
torch.cuda.set_device(0)
alltoall(1, 1)
sync()
got = dst[1][0 * cols:(0 + 1) * cols, :].cpu()          # columns of GPU 0 as received by GPU 1: [col][rows]
want = src[0][:, 1 * rows:2 * rows].cpu()
print("block check:", bool(torch.equal(got, want)))
