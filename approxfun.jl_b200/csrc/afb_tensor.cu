// 2-D tensor-product transform support (kernel K8 of SURVEY.md section 2c): tiled transpose and the
// multi-GPU (one process per GPU) slab decomposition with an NCCL all-to-all over NVLink.
//
// Replaces UPSTREAM ApproxFunBase transform!(::TensorSpace, M): factor-1 plan on every column, then
// factor-2 plan on every row (callers /root/reference/src/Plot/Plot.jl:290-332,
// src/Extras/sample.jl:196,218).
//
// NCCL is resolved lazily with dlopen so that the single-GPU library has no hard dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "afb_internal.h"

namespace afb {

// ---- transpose: in is column-major rows x cols (element (i,j) at i + j*rows); out is cols x rows ------
constexpr int TR_TILE = 32;
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                        int64_t rows, int64_t cols) {
    AFB_DYN_SMEM(double, tile);  // TR_TILE x (TR_TILE+1)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t i0 = (int64_t)blockIdx.x * TR_TILE, j0 = (int64_t)blockIdx.y * TR_TILE;
#pragma unroll
    for (int r = 0; r < TR_TILE; r += 8) {
        const int64_t i = i0 + tx, j = j0 + ty + r;
        if (i < rows && j < cols) tile[(ty + r) * (TR_TILE + 1) + tx] = in[i + j * rows];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR_TILE; r += 8) {
        const int64_t j = j0 + tx, i = i0 + ty + r;
        if (i < rows && j < cols) out[j + i * cols] = tile[tx * (TR_TILE + 1) + ty + r];
    }
}

int32_t launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, void* stream) {
    if (rows == 0 || cols == 0) return AFB_OK;
    const int64_t gx = (rows + TR_TILE - 1) / TR_TILE, gy = (cols + TR_TILE - 1) / TR_TILE;
    if (gy > 65535) return fail(AFB_UNSUPPORTED, "transpose: more than 2^21 columns");
    auto k = transpose_kernel;
    AFB_LAUNCH(k, dim3((unsigned)gx, (unsigned)gy), 256, sizeof(double) * TR_TILE * (TR_TILE + 1), stream, in, out, rows, cols);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ---- Padua transform helpers -----------------------------------------------------------
__global__ void scatter_idx_kernel(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ idx,
                                   int64_t N) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) dst[idx[p]] = src[p];
}
__global__ void gather_idx_kernel(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ idx,
                                  int64_t N, double scale) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) dst[p] = scale * src[idx[p]];
}
// idx[p] = iy + (d+2) ix on the (d+2) x (d+1) Chebyshev-Lobatto grid; x = -cospi(k/d) with k = d - ix
__global__ void paduapoints_kernel(double* __restrict__ xy, const int32_t* __restrict__ idx, int64_t N, int64_t d, double ax,
                                   double bx, double ay, double by) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int64_t ix = idx[p] / (d + 2), iy = idx[p] % (d + 2);
    const double x = -cospi((double)(d - ix) / (double)d);
    const double y = -cospi((double)(d + 1 - iy) / (double)(d + 1));
    xy[p] = 0.5 * (ax + bx) + 0.5 * (bx - ax) * x;
    xy[N + p] = 0.5 * (ay + by) + 0.5 * (by - ay) * y;
}
int32_t launch_scatter_idx(double* dst, const double* src, const int32_t* idx, int64_t N, void* stream) {
    if (N == 0) return AFB_OK;
    auto k = scatter_idx_kernel;
    AFB_LAUNCH(k, (unsigned)((N + 255) / 256), 256, 0, stream, dst, src, idx, N);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_gather_idx(double* dst, const double* src, const int32_t* idx, int64_t N, double scale, void* stream) {
    if (N == 0) return AFB_OK;
    auto k = gather_idx_kernel;
    AFB_LAUNCH(k, (unsigned)((N + 255) / 256), 256, 0, stream, dst, src, idx, N, scale);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_paduapoints(double* xy, const int32_t* idx, int64_t N, int64_t d, double ax, double bx, double ay, double by,
                           void* stream) {
    if (N == 0) return AFB_OK;
    auto k = paduapoints_kernel;
    AFB_LAUNCH(k, (unsigned)((N + 255) / 256), 256, 0, stream, xy, idx, N, d, ax, bx, ay, by);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// unpack after the all-to-all: R[s][i][jj] (s < G, i < nr, jj < nc) -> L[i][s*nc + jj]
__global__ void a2a_unpack_kernel(const double* __restrict__ R, double* __restrict__ L, int64_t G, int64_t nr, int64_t nc) {
    const int64_t total = G * nr * nc;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jj = g % nc, i = (g / nc) % nr, s = g / (nc * nr);
        L[i * (G * nc) + s * nc + jj] = R[g];
    }
}

// Fused transpose + all-to-all over NVLink peer memory.  A 32 (rows i) x 64 (columns j) tile of this rank's n x nc column
// block (after the column transforms) is transposed through shared memory and written STRAIGHT into the receive buffer of the
// rank that owns row i (local HBM for d == rank, a peer-mapped pointer otherwise): R_d[(i - d nr) * n2 + rank * nc + j], as
// 128-bit stores (a warp writes two 256-byte runs).  One kernel replaces pack-transpose + NCCL send/recv + unpack.
//   * Tiles are visited destination by destination, starting at rank + 1: at any moment the G ranks write to G different
//     destinations (no incast on one NVSwitch egress port).
//   * Completion is signalled on the device: every CTA fences its stores at system scope and counts itself on a local
//     counter; the last one stores the call's epoch into flags[rank] of EVERY destination (st.release.sys).  The row pass of
//     a destination is launched behind xchg_wait_kernel, which spins (ld.acquire.sys) until all G senders have reported the
//     epoch.  No host synchronisation and no NCCL on the data path.
//   * Receive buffers are double-buffered by epoch parity, so no "everyone is done reading" barrier is needed before the
//     writes: a sender reaches epoch e+2 only after the receiver's flag of epoch e+1, which the receiver stores after its
//     row pass of epoch e (stream order) -- see DESIGN.md section 6.
constexpr int XT_I = 32, XT_J = 64;
struct XchgPeers {
    double* recv[MAX_DEVICES];       // receive buffer of this epoch's parity on every rank
    uint64_t* flags[MAX_DEVICES];    // flags[d][s] = last epoch for which sender s's tiles have landed on rank d
};
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
#ifndef AFB_HOST_EMU
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
#ifndef AFB_HOST_EMU
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
// in: column-major n x nc (this rank's columns j0g..j0g+nc of the global matrix); rows [r0, r1) of every destination's
// nr-row share are sent (r0 = 0, r1 = nr: everything).  signal != 0: the last CTA publishes `epoch`.
__global__ void __launch_bounds__(256) xchg_transpose_kernel(const double* __restrict__ in, XchgPeers peers, int64_t n, int64_t nc,
                                                             int64_t nr, int64_t n2, int G, int rank, int64_t c0, int64_t c1,
                                                             int vec, unsigned* ctr, uint64_t epoch, int signal, int order) {
    AFB_DYN_SMEM(double, tile);  // XT_J x (XT_I + 1)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t ti_per = (nr + XT_I - 1) / XT_I, tj_cnt = (c1 - c0 + XT_J - 1) / XT_J;
    const int64_t per_dest = ti_per * tj_cnt;
    for (int64_t b = blockIdx.x; b < per_dest * G; b += gridDim.x) {
        // order 0: destination by destination, rotated (rank + 1 first, the local share last);
        // order 1: consecutive CTAs go to consecutive destinations (all peers in flight at once), rotated;
        // order 2: row tiles fastest over the whole matrix (unrotated)
        int k;
        int64_t w;
        if (order == 0) { k = (int)(b / per_dest); w = b - (int64_t)k * per_dest; }
        else if (order == 1) { k = (int)(b % G); w = b / G; }
        else if (order == 2) { const int64_t it = b % (ti_per * G); k = (int)(it / ti_per); w = (it % ti_per) + ti_per * (b / (ti_per * G)); }
        else { k = (int)(b / per_dest); w = b - (int64_t)k * per_dest; }   // order 3: as 0, column tiles fastest
        const int d = (order == 2) ? k : (rank + 1 + k) % G;
        // orders 0-2: consecutive CTAs walk down the rows of one column tile; order 3: along one band of 32 rows, so that the
        // CTAs in flight write one contiguous span of the destination (nc * 8 bytes per row)
        const int64_t il0 = (order == 3) ? (w / tj_cnt) * XT_I : (w % ti_per) * XT_I;
        const int64_t j0 = c0 + ((order == 3) ? (w % tj_cnt) : (w / ti_per)) * XT_J;
        const int64_t i0 = (int64_t)d * nr + il0;
#pragma unroll
        for (int r = 0; r < XT_J; r += 8) {
            const int64_t il = il0 + tx, j = j0 + ty + r;
            if (il < nr && j < c1) tile[(ty + r) * (XT_I + 1) + tx] = in[i0 + tx + j * n];
        }
        __syncthreads();
        double* dst = peers.recv[d] + (int64_t)rank * nc;
        // lanes 0-15: row ii, column pairs 0..15 ; lanes 16-31: row ii + 1
        const int half = tx >> 4, jp = tx & 15;
#pragma unroll
        for (int r = 0; r < XT_I; r += 16) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // column pairs jp and jp + 16... covers 64 columns in two halves
                const int ii = r + 2 * ty + half;
                const int jj = 2 * (jp + 16 * h);
                const int64_t il = il0 + ii, j = j0 + jj;
                if (il < nr && j < c1) {
                    const double a = tile[jj * (XT_I + 1) + ii];
                    if (vec && j + 1 < c1) {
                        const double bb = tile[(jj + 1) * (XT_I + 1) + ii];
                        *reinterpret_cast<double2*>(dst + il * n2 + j) = make_double2(a, bb);
                    } else {
                        dst[il * n2 + j] = a;
                        if (j + 1 < c1) dst[il * n2 + j + 1] = tile[(jj + 1) * (XT_I + 1) + ii];
                    }
                }
            }
        }
        __syncthreads();
    }
    if (!signal) return;
    // completion: fence this CTA's stores at system scope, count it; the last CTA publishes the epoch to every destination
    __syncthreads();
    if (threadIdx.x == 0) {
#ifndef AFB_HOST_EMU
        __threadfence_system();
        const unsigned prev = atomicAdd(ctr, 1u);
        if (prev == gridDim.x - 1) {
            *ctr = 0;                 // the next launch on this stream starts from zero
            __threadfence_system();
            for (int d = 0; d < G; ++d) st_release_sys(peers.flags[d] + rank, epoch);
        }
#endif
    }
}
// TMA variant of the exchange: a 32 (rows) x 128 (columns) tile is transposed into shared memory as 32 contiguous rows of
// 1 KB and every row leaves as ONE bulk copy (cp.async.bulk shared -> global, SASS UBLKCP.G.S) to the owner's receive
// buffer -- whole-packet NVLink writes issued by the copy engine of the SM instead of 16-byte LSU stores.  Two tile buffers:
// while the rows of tile k drain, tile k+1 is loaded and transposed.  Needs nr % 32 == 0, nc % 128 == 0, even n2.
constexpr int XB_I = 32, XB_J = 128, XB_LD = XB_J + 2;   // row pitch 1040 B: 16-byte aligned rows
__global__ void __launch_bounds__(256) xchg_bulk_kernel(const double* __restrict__ in, XchgPeers peers, int64_t n, int64_t nc,
                                                        int64_t nr, int64_t n2, int G, int rank, int64_t c0, int64_t c1,
                                                        unsigned* ctr, uint64_t epoch, int signal) {
    AFB_DYN_SMEM(double, tiles);  // 2 x XB_I x XB_LD
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t ti_per = nr / XB_I, tj_cnt = (c1 - c0) / XB_J;
    const int64_t per_dest = ti_per * tj_cnt;
    int buf = 0;
    for (int64_t b = blockIdx.x; b < per_dest * G; b += gridDim.x, buf ^= 1) {
        const int k = (int)(b / per_dest);
        const int64_t w = b - (int64_t)k * per_dest;
        const int d = (rank + 1 + k) % G;
        const int64_t il0 = (w / tj_cnt) * XB_I, j0 = c0 + (w % tj_cnt) * XB_J;   // column tiles fastest: contiguous destination
        const int64_t i0 = (int64_t)d * nr + il0;
        double v[XB_J / 8];
#pragma unroll
        for (int r = 0; r < XB_J / 8; ++r) v[r] = in[i0 + tx + (j0 + ty + 8 * r) * n];
        // the rows that last used this buffer (two tiles ago) must have been read out by the copy engine
        if (threadIdx.x < XB_I) bulk_wait_group_read<1>();
        __syncthreads();
        double* t = tiles + (size_t)buf * XB_I * XB_LD;
#pragma unroll
        for (int r = 0; r < XB_J / 8; ++r) t[tx * XB_LD + ty + 8 * r] = v[r];
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x < XB_I) {
            double* dst = peers.recv[d] + (il0 + threadIdx.x) * n2 + (int64_t)rank * nc + j0;
            bulk_copy_s2g(dst, t + threadIdx.x * XB_LD, (uint32_t)(sizeof(double) * XB_J));
            bulk_commit_group();
        }
    }
    if (threadIdx.x < XB_I) bulk_wait_group<0>();   // every row of this CTA has been written
    if (!signal) return;
    __syncthreads();
    if (threadIdx.x == 0) {
#ifndef AFB_HOST_EMU
        __threadfence_system();
        const unsigned prev = atomicAdd(ctr, 1u);
        if (prev == gridDim.x - 1) {
            *ctr = 0;
            __threadfence_system();
            for (int dd = 0; dd < G; ++dd) st_release_sys(peers.flags[dd] + rank, epoch);
        }
#endif
    }
}
// One warp: lane s < G waits until sender s has published `epoch` (or a later one).  A peer that never arrives would hang
// the stream; after ~10 s the kernel traps instead, which surfaces as a CUDA error at the caller's next synchronisation.
__global__ void xchg_wait_kernel(const uint64_t* __restrict__ flags, int G, uint64_t epoch) {
#ifndef AFB_HOST_EMU
    const int s = threadIdx.x;
    if (s >= G) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + s) < epoch) {
        __nanosleep(200);
        if (clock64() - t0 > 20000000000ll) __trap();
    }
#endif
}

}  // namespace afb

using namespace afb;

#ifndef AFB_HOST_EMU
// ---- minimal NCCL surface, resolved at run time -------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;
static std::mutex g_comm_mu;

static int32_t nccl_load() {
    if (g_nccl.h) return AFB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail(AFB_NCCL_ERROR, std::string("cannot load libnccl: ") + dlerror());
#define AFB_SYM(field, name)                                                     \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                            \
    if (!g_nccl.field) return fail(AFB_NCCL_ERROR, "libnccl lacks symbol " name)
    AFB_SYM(GetUniqueId, "ncclGetUniqueId");
    AFB_SYM(CommInitRank, "ncclCommInitRank");
    AFB_SYM(CommDestroy, "ncclCommDestroy");
    AFB_SYM(GroupStart, "ncclGroupStart");
    AFB_SYM(GroupEnd, "ncclGroupEnd");
    AFB_SYM(Send, "ncclSend");
    AFB_SYM(Recv, "ncclRecv");
    AFB_SYM(GetErrorString, "ncclGetErrorString");
#undef AFB_SYM
    return AFB_OK;
}
#define AFB_NCCL(expr)                                                                                      \
    do {                                                                                                    \
        ncclResult_t r_ = (expr);                                                                           \
        if (r_ != 0) return fail(AFB_NCCL_ERROR, std::string("NCCL: ") + g_nccl.GetErrorString(r_) + " in " #expr); \
    } while (0)
#endif

// One rank's share of the slab-decomposed 2-D transform (used by the one-process-per-GPU entry points afb_sharded2d_* and by
// the in-process multi-device plan behind afb_plan_execute_host).
struct Sharded2D {
    int32_t kind = 0;
    int64_t n = 0, n2 = 0, G = 1, rank = 0;
    int dev = 0;
    afb_plan col_plan = nullptr;  // length n,  batch n2/G
    afb_plan row_plan = nullptr;  // length n2, batch n/G
    double *tbuf = nullptr, *rbuf = nullptr;
    // peer-memory mode: ONE allocation [recv 0 | recv 1 | flags | counter] (one CUDA IPC handle in the multi-process case)
    char* xblock = nullptr;
    double* peer_recv[2][MAX_DEVICES] = {{nullptr}};
    uint64_t* peer_flags[MAX_DEVICES] = {nullptr};
    void* peer_base[MAX_DEVICES] = {nullptr};  // IPC mappings to close
    bool p2p = false, ipc = false;
    uint64_t epoch = 0;
    int chunks = 1;               // column-pass / exchange pipeline depth (see slab_execute)
    // exchange kernel tuning (afb_sharded2d_set_tuning).  Measured at 8 GPUs, 16384^2 (exchange alone, 235 MB per rank each way):
    // order 0 0.49 ms, 1 0.73, 2 0.72, 3 0.39 (8 CTAs per SM), 4 = TMA bulk rows 0.383 (3 CTAs per SM; 0.40 with one)
    int order = 4, grid_mult = 3, novec = 0;
    int phases = 7;               // measurement only: bit 0 column pass, bit 1 exchange + wait, bit 2 row pass
    cudaStream_t side = nullptr;  // the exchange runs here while the next column chunk is transformed
    cudaEvent_t ev_col[8] = {nullptr}, ev_x = nullptr;
    size_t recv_bytes() const { return sizeof(double) * (size_t)((n / G) * n2); }
    size_t xblock_bytes() const { return 2 * recv_bytes() + 4096; }
    double* recv(int parity) const { return reinterpret_cast<double*>(xblock + (size_t)parity * recv_bytes()); }
    uint64_t* flags() const { return reinterpret_cast<uint64_t*>(xblock + 2 * recv_bytes()); }
    unsigned* ctr() const { return reinterpret_cast<unsigned*>(xblock + 2 * recv_bytes() + 2048); }
};

static int32_t slab_alloc_xblock(Sharded2D& s) {
    if (s.xblock) return AFB_OK;
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, s.xblock_bytes()));
    s.xblock = reinterpret_cast<char*>(d);
    AFB_CUDA(cudaMemset(s.xblock + 2 * s.recv_bytes(), 0, 4096));
    return AFB_OK;
}
static void slab_wire_peer(Sharded2D& s, int d, char* base) {
    s.peer_recv[0][d] = reinterpret_cast<double*>(base);
    s.peer_recv[1][d] = reinterpret_cast<double*>(base + s.recv_bytes());
    s.peer_flags[d] = reinterpret_cast<uint64_t*>(base + 2 * s.recv_bytes());
}
static int32_t slab_create(Sharded2D& s, int32_t kind, int64_t n, int64_t n2, int64_t G, int64_t rank) {
    s.kind = kind; s.n = n; s.n2 = n2; s.G = G; s.rank = rank; s.dev = current_device();
    const int32_t k1 = (kind == AFB_CHEB1_2D_FWD) ? AFB_CHEB1_FWD : AFB_CHEB1_INV;
    AFB_TRY(afb_plan_create(&s.col_plan, k1, n, 0, n2 / G, n, 0));
    AFB_TRY(afb_plan_create(&s.row_plan, k1, n2, 0, n / G, n2, 0));
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)(n * (n2 / G))));
    s.rbuf = reinterpret_cast<double*>(d);
    return AFB_OK;
}
static void slab_free(Sharded2D& s) {
    if (s.col_plan) afb_plan_destroy(s.col_plan);
    if (s.row_plan) afb_plan_destroy(s.row_plan);
    if (s.tbuf) cudaFree(s.tbuf);
    if (s.rbuf) cudaFree(s.rbuf);
#ifndef AFB_HOST_EMU
    if (s.ipc)
        for (int64_t d = 0; d < s.G; ++d)
            if (d != s.rank && s.peer_base[d]) cudaIpcCloseMemHandle(s.peer_base[d]);
    for (auto& e : s.ev_col) if (e) cudaEventDestroy(e);
    if (s.ev_x) cudaEventDestroy(s.ev_x);
    if (s.side) cudaStreamDestroy(s.side);
#endif
    if (s.xblock) cudaFree(s.xblock);
    s = Sharded2D();
}

#ifndef AFB_HOST_EMU
static int32_t launch_xchg_bulk(Sharded2D& s, const XchgPeers& pp, int64_t c0, int64_t c1, uint64_t epoch, int signal, cudaStream_t st) {
    const int64_t G = s.G, nc = s.n2 / G, nr = s.n / G;
    const int64_t tiles = G * (nr / XB_I) * ((c1 - c0) / XB_J);
    const size_t smem = sizeof(double) * 2 * XB_I * XB_LD;
    const int64_t grid = std::min<int64_t>(tiles, std::min(s.grid_mult, 3) * (int64_t)sm_count());
    auto k = xchg_bulk_kernel;
    AFB_SMEM_OPTIN(k, smem);
    AFB_LAUNCH(k, (unsigned)grid, 256, smem, st, s.rbuf, pp, s.n, nc, nr, s.n2, (int)G, (int)s.rank, c0, c1, s.ctr(), epoch, signal);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
#endif

// Peer-memory execution of one rank: column transforms -> fused transpose straight into the owners' receive buffers ->
// wait for every sender's epoch flag -> row transforms.  With chunks > 1 the column pass runs in `chunks` pieces on the
// caller's stream and each finished piece is sent on a side stream while the next is transformed.
static int32_t slab_execute_p2p(Sharded2D& s, const double* d_in, double* d_out, void* stream) {
#ifdef AFB_HOST_EMU
    (void)s; (void)d_in; (void)d_out; (void)stream;
    return fail(AFB_UNSUPPORTED, "no peer memory in the emulator build");
#else
    const int64_t n = s.n, G = s.G, nc = s.n2 / G, nr = n / G;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t epoch = ++s.epoch;
    const int par = (int)(epoch & 1);
    XchgPeers pp;
    for (int d = 0; d < MAX_DEVICES; ++d) { pp.recv[d] = s.peer_recv[par][d]; pp.flags[d] = s.peer_flags[d]; }
    auto kt = xchg_transpose_kernel;
    const size_t smem = sizeof(double) * XT_J * (XT_I + 1);
    // chunk boundaries on multiples of 64 columns (vector stores stay aligned)
    int K = s.chunks;
    while (K > 1 && (nc / K) % XT_J) --K;
    if (K > 8) K = 8;
    if (s.phases != 7) {   // measurement of the single phases (results are meaningless unless all three run)
        if (s.phases & 1) AFB_TRY(afb_plan_execute_device(s.col_plan, d_in, s.rbuf, stream));
        if (s.phases & 2) {
            const int64_t tiles = G * ((nr + XT_I - 1) / XT_I) * ((nc + XT_J - 1) / XT_J);
            if (s.order == 4 && nr % XB_I == 0 && nc % XB_J == 0 && s.n2 % 2 == 0) {
                AFB_TRY(launch_xchg_bulk(s, pp, 0, nc, epoch, 1, st));
            } else {
                const int64_t grid = std::min<int64_t>(tiles, (s.order == 4 ? 8 : s.grid_mult) * (int64_t)sm_count());
                const int vec = (!s.novec && nc % 2 == 0 && s.n2 % 2 == 0) ? 1 : 0;
                AFB_LAUNCH(kt, (unsigned)grid, 256, smem, st, s.rbuf, pp, n, nc, nr, s.n2, (int)G, (int)s.rank, (int64_t)0, nc, vec,
                           s.ctr(), epoch, 1, s.order == 4 ? 3 : s.order);
                AFB_KERNEL_CHECK();
            }
            auto kw0 = xchg_wait_kernel;
            AFB_LAUNCH(kw0, 1, 32, 0, st, s.flags(), (int)G, epoch);
            AFB_KERNEL_CHECK();
        }
        if (s.phases & 4) AFB_TRY(afb_plan_execute_device(s.row_plan, s.recv(par), d_out, stream));
        return AFB_OK;
    }
    if (K == 1) {
        AFB_TRY(afb_plan_execute_device(s.col_plan, d_in, s.rbuf, stream));
        const int64_t tiles = G * ((nr + XT_I - 1) / XT_I) * ((nc + XT_J - 1) / XT_J);
        if (s.order == 4 && nr % XB_I == 0 && nc % XB_J == 0 && s.n2 % 2 == 0) {
            AFB_TRY(launch_xchg_bulk(s, pp, 0, nc, epoch, 1, st));
        } else {
            const int64_t grid = std::min<int64_t>(tiles, (s.order == 4 ? 8 : s.grid_mult) * (int64_t)sm_count());
            const int vec = (!s.novec && nc % 2 == 0 && s.n2 % 2 == 0) ? 1 : 0;
            AFB_LAUNCH(kt, (unsigned)grid, 256, smem, st, s.rbuf, pp, n, nc, nr, s.n2, (int)G, (int)s.rank, (int64_t)0, nc, vec,
                       s.ctr(), epoch, 1, s.order == 4 ? 3 : s.order);
            AFB_KERNEL_CHECK();
        }
    } else {
        if (!s.side) AFB_CUDA(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
        if (!s.ev_x) AFB_CUDA(cudaEventCreateWithFlags(&s.ev_x, cudaEventDisableTiming));
        const int64_t cw = nc / K;
        for (int c = 0; c < K; ++c) {
            if (!s.ev_col[c]) AFB_CUDA(cudaEventCreateWithFlags(&s.ev_col[c], cudaEventDisableTiming));
            AFB_TRY(slab_col_chunk(s.col_plan, d_in + c * cw * n, s.rbuf + c * cw * n, cw, stream));
            AFB_CUDA(cudaEventRecord(s.ev_col[c], st));
            AFB_CUDA(cudaStreamWaitEvent(s.side, s.ev_col[c], 0));
            const int64_t tiles = G * ((nr + XT_I - 1) / XT_I) * (cw / XT_J);
            // a few CTAs per chunk: the exchange is NVLink-bound and must leave the SMs to the next column chunk
            const int64_t grid = std::min<int64_t>(tiles, c + 1 < K ? 2 * (int64_t)sm_count() : 8 * (int64_t)sm_count());
            AFB_LAUNCH(kt, (unsigned)grid, 256, smem, s.side, s.rbuf, pp, n, nc, nr, s.n2, (int)G, (int)s.rank, c * cw, (c + 1) * cw, 1,
                       s.ctr(), epoch, c + 1 == K ? 1 : 0, s.order == 4 ? 3 : s.order);
            AFB_KERNEL_CHECK();
        }
        AFB_CUDA(cudaEventRecord(s.ev_x, s.side));
        AFB_CUDA(cudaStreamWaitEvent(st, s.ev_x, 0));
    }
    auto kw = xchg_wait_kernel;
    AFB_LAUNCH(kw, 1, 32, 0, st, s.flags(), (int)G, epoch);
    AFB_KERNEL_CHECK();
    return afb_plan_execute_device(s.row_plan, s.recv(par), d_out, stream);
#endif
}

extern "C" {

int32_t afb_comm_unique_id(void* id128) {
#ifdef AFB_HOST_EMU
    (void)id128;
    return fail(AFB_UNSUPPORTED, "no NCCL in the emulator build");
#else
    if (!id128) return fail(AFB_BAD_ARG, "afb_comm_unique_id: null buffer");
    AFB_TRY(nccl_load());
    ncclUniqueId id;
    AFB_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return AFB_OK;
#endif
}

int32_t afb_comm_create(const void* id128, int32_t rank, int32_t world) {
#ifdef AFB_HOST_EMU
    (void)id128; (void)rank; (void)world;
    return fail(AFB_UNSUPPORTED, "no NCCL in the emulator build");
#else
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(AFB_BAD_ARG, "afb_comm_create: bad arguments");
    AFB_TRY(ensure_device());
    AFB_TRY(nccl_load());
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm) return fail(AFB_BAD_ARG, "afb_comm_create: communicator already exists");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    AFB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    g_rank = rank;
    g_world = world;
    return AFB_OK;
#endif
}

int32_t afb_comm_destroy(void) {
#ifndef AFB_HOST_EMU
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm) {
        AFB_NCCL(g_nccl.CommDestroy(g_comm));
        g_comm = nullptr;
    }
#endif
    return AFB_OK;
}

}  // extern "C"

// The sharded entry points use their own handle type (afb_plan_s stays private to afb_api.cu).
namespace {
struct ShardHandle {
    uint64_t magic = 0x5348415244324421ull;  // "SHARD2D!"
    Sharded2D s;
};
}  // namespace

extern "C" {

int32_t afb_sharded2d_create(void** out, int32_t kind, int64_t n, int64_t n2) {
#ifdef AFB_HOST_EMU
    (void)out; (void)kind; (void)n; (void)n2;
    return fail(AFB_UNSUPPORTED, "no NCCL in the emulator build");
#else
    if (!out) return fail(AFB_BAD_ARG, "afb_sharded2d_create: null out");
    if (kind != AFB_CHEB1_2D_FWD && kind != AFB_CHEB1_2D_INV) return fail(AFB_BAD_ARG, "afb_sharded2d_create: kind");
    if (!g_comm) return fail(AFB_BAD_ARG, "afb_sharded2d_create: call afb_comm_create first");
    const int64_t G = g_world;
    if (n <= 0 || n2 <= 0 || n % G || n2 % G) return fail(AFB_BAD_SIZE, "afb_sharded2d_create: n and n2 must be multiples of the world size");
    ShardHandle* h = new ShardHandle();
    int32_t st = slab_create(h->s, kind, n, n2, G, g_rank);
    if (st == AFB_OK) {  // pack buffer of the NCCL baseline route
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, sizeof(double) * (size_t)(n * (n2 / G)));
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
        h->s.tbuf = reinterpret_cast<double*>(d);
    }
    if (st != AFB_OK) { slab_free(h->s); delete h; return st; }
    *out = h;
    return AFB_OK;
#endif
}

// Peer-memory mode (one process per GPU): each rank exports the CUDA IPC handle of its receive block
// (64 bytes), the host layer all-gathers them, and every rank maps its peers' blocks.
int32_t afb_sharded2d_ipc_export(void* handle, void* handle64) {
#ifdef AFB_HOST_EMU
    (void)handle; (void)handle64;
    return fail(AFB_UNSUPPORTED, "no CUDA IPC in the emulator build");
#else
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || !handle64) return fail(AFB_BAD_ARG, "afb_sharded2d_ipc_export: bad argument");
    Sharded2D& s = h->s;
    AFB_TRY(slab_alloc_xblock(s));
    cudaIpcMemHandle_t mh;
    AFB_CUDA(cudaIpcGetMemHandle(&mh, s.xblock));
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &mh, 64);
    return AFB_OK;
#endif
}

int32_t afb_sharded2d_ipc_open(void* handle, const void* handles, int32_t world) {
#ifdef AFB_HOST_EMU
    (void)handle; (void)handles; (void)world;
    return fail(AFB_UNSUPPORTED, "no CUDA IPC in the emulator build");
#else
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || !handles) return fail(AFB_BAD_ARG, "afb_sharded2d_ipc_open: bad argument");
    Sharded2D& s = h->s;
    if (world != s.G || world > MAX_DEVICES || !s.xblock) return fail(AFB_BAD_ARG, "afb_sharded2d_ipc_open: export first; world <= 16");
    for (int64_t d = 0; d < s.G; ++d) {
        if (d == s.rank) { slab_wire_peer(s, (int)d, s.xblock); continue; }
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, reinterpret_cast<const char*>(handles) + 64 * d, 64);
        void* p = nullptr;
        AFB_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        s.peer_base[d] = p;
        slab_wire_peer(s, (int)d, reinterpret_cast<char*>(p));
    }
    s.p2p = true;
    s.ipc = true;
    return AFB_OK;
#endif
}

int32_t afb_sharded2d_set_chunks(void* handle, int32_t chunks) {
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || chunks < 1 || chunks > 8) return fail(AFB_BAD_ARG, "afb_sharded2d_set_chunks: bad argument");
    h->s.chunks = chunks;
    return AFB_OK;
}

int32_t afb_sharded2d_set_tuning(void* handle, int32_t order, int32_t grid_mult, int32_t novec) {
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || order < 0 || order > 4 || grid_mult < 1 || grid_mult > 64)
        return fail(AFB_BAD_ARG, "afb_sharded2d_set_tuning: bad argument");
    h->s.order = order; h->s.grid_mult = grid_mult; h->s.novec = novec ? 1 : 0;
    return AFB_OK;
}
int32_t afb_sharded2d_set_phases(void* handle, int32_t mask) {
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || mask < 1 || mask > 7) return fail(AFB_BAD_ARG, "afb_sharded2d_set_phases: bad argument");
    h->s.phases = mask;
    return AFB_OK;
}

// d_in : this rank's column block, column-major n x (n2/G)
// d_out: this rank's ROW block of the result, stored as n/G contiguous lines of length n2
//        (= the transposed coefficient matrix sharded by columns); see DESIGN.md "multi-GPU 2-D"
int32_t afb_sharded2d_execute_device(void* handle, const double* d_in, double* d_out, void* stream) {
#ifdef AFB_HOST_EMU
    (void)handle; (void)d_in; (void)d_out; (void)stream;
    return fail(AFB_UNSUPPORTED, "no NCCL in the emulator build");
#else
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull) return fail(AFB_BAD_ARG, "afb_sharded2d_execute_device: bad handle");
    Sharded2D& s = h->s;
    if (s.p2p) return slab_execute_p2p(s, d_in, d_out, stream);
    const int64_t n = s.n, G = s.G, nc = s.n2 / G, nr = n / G;
    // NCCL baseline route.  1. local column transforms (into rbuf to keep d_in intact)
    AFB_TRY(afb_plan_execute_device(s.col_plan, d_in, s.rbuf, stream));
    // 2. pack = local transpose: T[i*nc + j], rows of destination d are the contiguous range i in [d nr, (d+1) nr)
    AFB_TRY(launch_transpose(s.rbuf, s.tbuf, n, nc, stream));
    // 3. all-to-all over NVLink (grouped send/recv: the installed NCCL header has no ncclAlltoAll)
    AFB_NCCL(g_nccl.GroupStart());
    for (int64_t d = 0; d < G; ++d) {
        AFB_NCCL(g_nccl.Send(s.tbuf + d * nr * nc, (size_t)(nr * nc), ncclFloat64, (int)d, g_comm, (cudaStream_t)stream));
        AFB_NCCL(g_nccl.Recv(s.rbuf + d * nr * nc, (size_t)(nr * nc), ncclFloat64, (int)d, g_comm, (cudaStream_t)stream));
    }
    AFB_NCCL(g_nccl.GroupEnd());
    // 4. unpack to full row lines, 5. local row transforms
    auto k = a2a_unpack_kernel;
    const int64_t total = G * nr * nc;
    int64_t grid = std::min<int64_t>((total + 255) / 256, 16 * (int64_t)sm_count());
    AFB_LAUNCH(k, (unsigned)grid, 256, 0, stream, s.rbuf, d_out, G, nr, nc);
    AFB_KERNEL_CHECK();
    return afb_plan_execute_device(s.row_plan, d_out, d_out, stream);
#endif
}

int32_t afb_sharded2d_destroy(void* handle) {
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h) return AFB_OK;
    if (h->magic != 0x5348415244324421ull) return fail(AFB_BAD_ARG, "afb_sharded2d_destroy: bad handle");
    slab_free(h->s);
    h->magic = 0;
    delete h;
    return AFB_OK;
}

}  // extern "C"

// =====================================================================================================================
// In-process multi-device 2-D plan (afb_init(devices, ndev) with ndev > 1; reached through afb_plan_execute_host): the same
// slab decomposition with every device driven by one host thread of THIS process.  Peer buffers are plain pointers
// (cudaDeviceEnablePeerAccess in afb_init): no IPC, no NCCL, no torch.
// =====================================================================================================================
namespace afb {

struct MultiDev2D {
    std::vector<int> devs;
    std::vector<Sharded2D> slab;
    std::vector<double*> din, dout, tout;   // per device: column block, row lines, row block transposed back to column-major
    std::vector<cudaStream_t> stream;
    int64_t n = 0, n2 = 0;
};

bool mdev2d_supported(int32_t kind, int64_t n, int64_t n2, int ndev, uint32_t flags) {
    (void)flags;
    if (kind != AFB_CHEB1_2D_FWD && kind != AFB_CHEB1_2D_INV) return false;
    return ndev > 1 && ndev <= MAX_DEVICES && n > 0 && n2 > 0 && n % ndev == 0 && n2 % ndev == 0 && n * n2 >= ((int64_t)1 << 21);
}

bool mdev2d_matches(void* handle, const std::vector<int>& devs) {
    MultiDev2D* m = reinterpret_cast<MultiDev2D*>(handle);
    return m && m->devs == devs;
}

void mdev2d_destroy(void* handle) {
    MultiDev2D* m = reinterpret_cast<MultiDev2D*>(handle);
    if (!m) return;
    int keep = 0;
    cudaGetDevice(&keep);
    for (size_t g = 0; g < m->slab.size(); ++g) {
        cudaSetDevice(m->devs[g]);
        slab_free(m->slab[g]);
        if (g < m->din.size() && m->din[g]) cudaFree(m->din[g]);
        if (g < m->dout.size() && m->dout[g]) cudaFree(m->dout[g]);
        if (g < m->tout.size() && m->tout[g]) cudaFree(m->tout[g]);
        if (g < m->stream.size() && m->stream[g]) cudaStreamDestroy(m->stream[g]);
    }
    cudaSetDevice(keep);
    delete m;
}

int32_t mdev2d_create(void** out, int32_t kind, int64_t n, int64_t n2, const std::vector<int>& devs) {
#ifdef AFB_HOST_EMU
    (void)out; (void)kind; (void)n; (void)n2; (void)devs;
    return fail(AFB_UNSUPPORTED, "no peer memory in the emulator build");
#else
    const int G = (int)devs.size();
    MultiDev2D* m = new MultiDev2D();
    m->devs = devs; m->n = n; m->n2 = n2;
    m->slab.resize((size_t)G);
    m->din.assign((size_t)G, nullptr); m->dout.assign((size_t)G, nullptr); m->tout.assign((size_t)G, nullptr);
    m->stream.assign((size_t)G, nullptr);
    int keep = 0;
    cudaGetDevice(&keep);
    int32_t st = AFB_OK;
    const size_t blk = sizeof(double) * (size_t)(n * (n2 / G));
    for (int g = 0; g < G && st == AFB_OK; ++g) {
        cudaError_t e = cudaSetDevice(devs[(size_t)g]);
        if (e != cudaSuccess) { st = cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); break; }
        st = ensure_device();
        if (st == AFB_OK) st = slab_create(m->slab[(size_t)g], kind, n, n2, G, g);
        if (st == AFB_OK) st = slab_alloc_xblock(m->slab[(size_t)g]);
        for (double** b : {&m->din[(size_t)g], &m->dout[(size_t)g], &m->tout[(size_t)g]}) {
            if (st != AFB_OK) break;
            void* d = nullptr;
            e = cudaMalloc(&d, blk);
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(2-D slab)", __FILE__, __LINE__);
            *b = reinterpret_cast<double*>(d);
        }
        if (st == AFB_OK) {
            e = cudaStreamCreateWithFlags(&m->stream[(size_t)g], cudaStreamNonBlocking);
            if (e != cudaSuccess) st = cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
        }
    }
    if (st == AFB_OK) {
        for (int g = 0; g < G; ++g) {
            for (int d = 0; d < G; ++d) slab_wire_peer(m->slab[(size_t)g], d, m->slab[(size_t)d].xblock);
            m->slab[(size_t)g].p2p = true;
        }
    }
    cudaSetDevice(keep);
    if (st != AFB_OK) { mdev2d_destroy(m); return st; }
    *out = m;
    return AFB_OK;
#endif
}

// in / out: the whole column-major n x n2 matrix on the host.  Device g takes columns [g nc, (g+1) nc) (one contiguous
// block), and after the exchange owns rows [g nr, (g+1) nr) as lines; they are transposed back on the device and land in
// `out` through a strided (pitch n) device-to-host copy.
int32_t mdev2d_execute_host(void* handle, const double* in, double* out) {
#ifdef AFB_HOST_EMU
    (void)handle; (void)in; (void)out;
    return fail(AFB_UNSUPPORTED, "no peer memory in the emulator build");
#else
    MultiDev2D* m = reinterpret_cast<MultiDev2D*>(handle);
    const int G = (int)m->devs.size();
    const int64_t n = m->n, n2 = m->n2, nc = n2 / G, nr = n / G;
    std::vector<int32_t> st((size_t)G, AFB_OK);
    std::vector<std::string> msg((size_t)G);
    std::vector<std::thread> th;
    int keep = 0;
    AFB_CUDA(cudaGetDevice(&keep));
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            auto run = [&]() -> int32_t {
                AFB_CUDA(cudaSetDevice(m->devs[(size_t)g]));
                cudaStream_t s = m->stream[(size_t)g];
                // pageable caller arrays go through the library's pinned ring (a pageable cudaMemcpyAsync is staged by the
                // driver on this thread: 2 GPUs 189 ms against 56 ms for registered arrays at 16 384^2)
                if (host_is_pageable(in)) {
                    AFB_TRY(host_big_h2d(m->din[(size_t)g], in + (size_t)g * nc * n, sizeof(double) * (size_t)(n * nc)));
                } else {
                    AFB_CUDA(cudaMemcpyAsync(m->din[(size_t)g], in + (size_t)g * nc * n, sizeof(double) * (size_t)(n * nc),
                                             cudaMemcpyHostToDevice, s));
                }
                AFB_TRY(slab_execute_p2p(m->slab[(size_t)g], m->din[(size_t)g], m->dout[(size_t)g], s));
                AFB_TRY(launch_transpose(m->dout[(size_t)g], m->tout[(size_t)g], n2, nr, s));
                if (host_is_pageable(out)) {
                    AFB_CUDA(cudaStreamSynchronize(s));
                    AFB_TRY(host_big_d2h_2d(out + (size_t)g * nr, sizeof(double) * (size_t)n, m->tout[(size_t)g],
                                            sizeof(double) * (size_t)nr, sizeof(double) * (size_t)nr, (size_t)n2));
                } else {
                    AFB_CUDA(cudaMemcpy2DAsync(out + (size_t)g * nr, sizeof(double) * (size_t)n, m->tout[(size_t)g],
                                               sizeof(double) * (size_t)nr, sizeof(double) * (size_t)nr, (size_t)n2,
                                               cudaMemcpyDeviceToHost, s));
                }
                AFB_CUDA(cudaStreamSynchronize(s));
                return AFB_OK;
            };
            st[(size_t)g] = run();
            if (st[(size_t)g] != AFB_OK) { char buf[512]; afb_last_error(buf, sizeof buf); msg[(size_t)g] = buf; }
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(keep);
    for (int g = 0; g < G; ++g)
        if (st[(size_t)g] != AFB_OK) return fail(st[(size_t)g], msg[(size_t)g]);
    return AFB_OK;
#endif
}

}  // namespace afb
