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

// Fused transpose + all-to-all over NVLink peer memory: tile (i, j) of this rank's n x nc column block (after the
// column transforms) is transposed through shared memory and written STRAIGHT into the row-line buffer of the rank
// that owns row i (local HBM for d == rank, a peer-mapped pointer otherwise): L_d[(i - d nr) * n2 + rank * nc + j].
// One kernel replaces pack-transpose + NCCL send/recv + unpack; stores are 256-byte contiguous runs per row.
struct PeerPtrs { double* p[16]; };
__global__ void __launch_bounds__(256) transpose_to_peers_kernel(const double* __restrict__ in, PeerPtrs peers, int64_t n,
                                                                 int64_t nc, int64_t nr, int64_t n2, int64_t rank) {
    AFB_DYN_SMEM(double, tile);  // TR_TILE x (TR_TILE+1)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t i0 = (int64_t)blockIdx.x * TR_TILE, j0 = (int64_t)blockIdx.y * TR_TILE;
#pragma unroll
    for (int r = 0; r < TR_TILE; r += 8) {
        const int64_t i = i0 + tx, j = j0 + ty + r;
        if (i < n && j < nc) tile[(ty + r) * (TR_TILE + 1) + tx] = in[i + j * n];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR_TILE; r += 8) {
        const int64_t j = j0 + tx, i = i0 + ty + r;
        if (i < n && j < nc) {
            const int64_t d = i / nr;
            peers.p[d][(i - d * nr) * n2 + rank * nc + j] = tile[tx * (TR_TILE + 1) + ty + r];
        }
    }
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

struct Sharded2D {
    int32_t kind;
    int64_t n, n2, G, rank;
    afb_plan col_plan = nullptr;  // length n,  batch n2/G
    afb_plan row_plan = nullptr;  // length n2, batch n/G
    double *tbuf = nullptr, *rbuf = nullptr;
    // peer-memory (CUDA IPC) mode: lbuf receives every rank's tiles; peer[d] = rank d's lbuf mapped here
    double* lbuf = nullptr;
    double* peer[16] = {nullptr};
    bool p2p = false;
    double* sync_buf = nullptr;  // 2 x G doubles for the NCCL stream barrier
};

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

// The sharded plan is an ordinary afb_plan whose handle wraps Sharded2D; to keep afb_plan_s private to
// afb_api.cu the sharded entry points live here and use their own handle type behind the same typedef.
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
    h->s.kind = kind; h->s.n = n; h->s.n2 = n2; h->s.G = G; h->s.rank = g_rank;
    const int32_t k1 = (kind == AFB_CHEB1_2D_FWD) ? AFB_CHEB1_FWD : AFB_CHEB1_INV;
    int32_t st = afb_plan_create(&h->s.col_plan, k1, n, 0, n2 / G, n, 0);
    if (st == AFB_OK) st = afb_plan_create(&h->s.row_plan, k1, n2, 0, n / G, n2, 0);
    const size_t bytes = sizeof(double) * (size_t)(n * (n2 / G));
    if (st == AFB_OK) { void* d = nullptr; cudaError_t e = cudaMalloc(&d, bytes); if (e) st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); h->s.tbuf = (double*)d; }
    if (st == AFB_OK) { void* d = nullptr; cudaError_t e = cudaMalloc(&d, bytes); if (e) st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); h->s.rbuf = (double*)d; }
    if (st != AFB_OK) {
        if (h->s.col_plan) afb_plan_destroy(h->s.col_plan);
        if (h->s.row_plan) afb_plan_destroy(h->s.row_plan);
        if (h->s.tbuf) cudaFree(h->s.tbuf);
        if (h->s.rbuf) cudaFree(h->s.rbuf);
        delete h;
        return st;
    }
    *out = h;
    return AFB_OK;
#endif
}

#ifndef AFB_HOST_EMU
// stream-ordered barrier over the communicator (one double to and from every peer)
static int32_t nccl_stream_barrier(Sharded2D& s, void* stream) {
    AFB_NCCL(g_nccl.GroupStart());
    for (int64_t d = 0; d < s.G; ++d) {
        if (d == s.rank) continue;
        AFB_NCCL(g_nccl.Send(s.sync_buf + d, 1, ncclFloat64, (int)d, g_comm, (cudaStream_t)stream));
        AFB_NCCL(g_nccl.Recv(s.sync_buf + s.G + d, 1, ncclFloat64, (int)d, g_comm, (cudaStream_t)stream));
    }
    AFB_NCCL(g_nccl.GroupEnd());
    return AFB_OK;
}
#endif

// Peer-memory mode (one process per GPU): each rank exports the CUDA IPC handle of its receive buffer
// (64 bytes), the host layer all-gathers them, and every rank maps its peers' buffers.
int32_t afb_sharded2d_ipc_export(void* handle, void* handle64) {
#ifdef AFB_HOST_EMU
    (void)handle; (void)handle64;
    return fail(AFB_UNSUPPORTED, "no CUDA IPC in the emulator build");
#else
    ShardHandle* h = reinterpret_cast<ShardHandle*>(handle);
    if (!h || h->magic != 0x5348415244324421ull || !handle64) return fail(AFB_BAD_ARG, "afb_sharded2d_ipc_export: bad argument");
    Sharded2D& s = h->s;
    if (!s.lbuf) {
        void* d = nullptr;
        AFB_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)((s.n / s.G) * s.n2)));
        s.lbuf = reinterpret_cast<double*>(d);
    }
    cudaIpcMemHandle_t mh;
    AFB_CUDA(cudaIpcGetMemHandle(&mh, s.lbuf));
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
    if (world != s.G || world > 16 || !s.lbuf) return fail(AFB_BAD_ARG, "afb_sharded2d_ipc_open: export first; world <= 16");
    for (int64_t d = 0; d < s.G; ++d) {
        if (d == s.rank) { s.peer[d] = s.lbuf; continue; }
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, reinterpret_cast<const char*>(handles) + 64 * d, 64);
        void* p = nullptr;
        AFB_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        s.peer[d] = reinterpret_cast<double*>(p);
    }
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)(2 * s.G)));
    AFB_CUDA(cudaMemset(d, 0, sizeof(double) * (size_t)(2 * s.G)));
    s.sync_buf = reinterpret_cast<double*>(d);
    s.p2p = true;
    return AFB_OK;
#endif
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
    const int64_t n = s.n, G = s.G, nc = s.n2 / G, nr = n / G;
    // 1. local column transforms (in place into rbuf to keep d_in intact)
    AFB_TRY(afb_plan_execute_device(s.col_plan, d_in, s.rbuf, stream));
    if (s.p2p) {
        // every rank must be done reading its lbuf (previous call's row transforms) before anyone overwrites it
        AFB_TRY(nccl_stream_barrier(s, stream));
        PeerPtrs pp;
        for (int d = 0; d < 16; ++d) pp.p[d] = s.peer[d];
        const int64_t gx = (n + TR_TILE - 1) / TR_TILE, gy = (nc + TR_TILE - 1) / TR_TILE;
        auto kt = transpose_to_peers_kernel;
        AFB_LAUNCH(kt, dim3((unsigned)gx, (unsigned)gy), 256, sizeof(double) * TR_TILE * (TR_TILE + 1), stream, s.rbuf, pp, n, nc,
                   nr, s.n2, s.rank);
        AFB_KERNEL_CHECK();
        // all peers' tiles have landed in my lbuf once every rank has passed this barrier
        AFB_TRY(nccl_stream_barrier(s, stream));
        return afb_plan_execute_device(s.row_plan, s.lbuf, d_out, stream);
    }
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
    if (h->s.col_plan) afb_plan_destroy(h->s.col_plan);
    if (h->s.row_plan) afb_plan_destroy(h->s.row_plan);
    if (h->s.tbuf) cudaFree(h->s.tbuf);
    if (h->s.rbuf) cudaFree(h->s.rbuf);
#ifndef AFB_HOST_EMU
    for (int64_t d = 0; d < h->s.G; ++d)
        if (h->s.p2p && d != h->s.rank && h->s.peer[d]) cudaIpcCloseMemHandle(h->s.peer[d]);
#endif
    if (h->s.lbuf) cudaFree(h->s.lbuf);
    if (h->s.sync_buf) cudaFree(h->s.sync_buf);
    h->magic = 0;
    delete h;
    return AFB_OK;
}

}  // extern "C"
