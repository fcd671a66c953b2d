// extern "C" boundary of libapproxfun_b200.so: status plumbing, twiddle tables, plans, host wrappers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "afb_internal.h"

namespace afb {

// ---- status --------------------------------------------------------------------------------------
static thread_local std::string g_err_msg;
static thread_local int32_t g_err_code = AFB_OK;

void set_error(int32_t code, const std::string& msg) {
    g_err_code = code;
    g_err_msg = msg;
}
int32_t fail(int32_t code, const std::string& msg) {
    set_error(code, msg);
    return code;
}
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear sticky-less errors
    return fail(e == cudaErrorMemoryAllocation ? AFB_OOM : AFB_CUDA_ERROR, buf);
}

// ---- devices ----------------------------------------------------------------------------------------
// The library follows the calling thread's current CUDA device (as cuFFT does): tables, staging pools and the block cache
// are kept per device, a plan belongs to the device that was current when it was created, and afb_init(devices, ndev)
// names the set of devices over which the host-pointer entry points fan out (one host thread and one stream set each).
struct DeviceState {
    std::atomic<bool> ok{false};
    int sms = 0;
    std::mutex mu;                                   // tables
    std::map<int64_t, cplx*> root_tables, quarter_tables;
};
static DeviceState g_dev[MAX_DEVICES];
static std::mutex g_mu;                              // device list, first-touch of a device
static std::vector<int> g_devlist;                   // afb_init's device set (empty: the current device only)

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}

int32_t ensure_device() {
    int n = 0, dev = 0;
    // fast path: the current device has been seen before
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < MAX_DEVICES && g_dev[dev].ok.load(std::memory_order_acquire))
        return AFB_OK;
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_mu);
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(AFB_NO_DEVICE, "no CUDA device visible: libapproxfun_b200 has no CPU fallback");
    }
    AFB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(AFB_UNSUPPORTED, "device ordinal beyond MAX_DEVICES (16)");
    if (g_dev[dev].ok.load(std::memory_order_acquire)) return AFB_OK;
    cudaDeviceProp prop;
    AFB_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_dev[dev].sms = prop.multiProcessorCount;
    g_dev[dev].ok.store(true, std::memory_order_release);
    return AFB_OK;
}
int sm_count() {
    const int s = g_dev[current_device()].sms;
    return s > 0 ? s : 148;
}
const std::vector<int>& device_list() { return g_devlist; }

// ---- tables (per device) -------------------------------------------------------------------------------
int32_t upload_table(const std::vector<cplx>& host, cplx** dev) {
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, sizeof(cplx) * std::max<size_t>(host.size(), 1)));
    cudaError_t e = cudaMemcpy(d, host.data(), sizeof(cplx) * host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "cudaMemcpy(table)", __FILE__, __LINE__); }
    *dev = reinterpret_cast<cplx*>(d);
    return AFB_OK;
}

static inline cplx unit_ld(long double turns_num, long double turns_den, int sign) {
    // exp(sign * 2 pi i * num/den) with the angle reduced in extended precision
    const long double PI2 = 6.283185307179586476925286766559005768L;
    const long double a = PI2 * (turns_num / turns_den);
    return make_double2((double)cosl(a), (double)(sign * sinl(a)));
}

int32_t root_table(int64_t N, const cplx** out) {
    DeviceState& ds = g_dev[current_device()];
    std::lock_guard<std::mutex> lk(ds.mu);
    auto it = ds.root_tables.find(N);
    if (it != ds.root_tables.end()) { *out = it->second; return AFB_OK; }
    std::vector<cplx> h((size_t)N);
    for (int64_t m = 0; m < N; ++m) h[(size_t)m] = unit_ld((long double)m, (long double)N, -1);
    cplx* d = nullptr;
    AFB_TRY(upload_table(h, &d));
    ds.root_tables[N] = d;
    *out = d;
    return AFB_OK;
}

int32_t quarter_table(int64_t n, const cplx** out) {
    DeviceState& ds = g_dev[current_device()];
    std::lock_guard<std::mutex> lk(ds.mu);
    auto it = ds.quarter_tables.find(n);
    if (it != ds.quarter_tables.end()) { *out = it->second; return AFB_OK; }
    std::vector<cplx> h((size_t)n + 1);
    for (int64_t k = 0; k <= n; ++k) h[(size_t)k] = unit_ld((long double)k, (long double)(4 * n), -1);
    cplx* d = nullptr;
    AFB_TRY(upload_table(h, &d));
    ds.quarter_tables[n] = d;
    *out = d;
    return AFB_OK;
}

// Tables live as long as the process: plans (and the Python layer's plan cache) keep raw pointers into them, so
// afb_shutdown does NOT free them (a few MB per length); it only drops the staging pools and the block cache.
void free_tables() {}

}  // namespace afb

using namespace afb;

// Staging pipes of the host-pointer entry points: NPIPE streams with grow-only device buffers.  Every device keeps a
// free list of pipes; a call leases one (or makes a new one), so calls on different plans -- and on different devices --
// never wait for each other (include/approxfun_b200.h "Threading").
// Pageable host arrays (a Julia Vector{Float64}, a NumPy array) cannot be DMA-ed directly: the driver would stage them through
// its own bounce buffers with ONE copying thread (measured: 65536 x 4096 doubles 317 ms per call against 45 ms for pinned
// buffers).  The library stages them itself: a small team of host threads copies chunk c+1 into a pinned ring slot (and chunk
// c-2 out of one) while the GPU and the DMA engines work on the chunks in between.
class CopyTeam {
  public:
    // dst / src rows of `width` bytes, `height` rows, pitches in bytes; returns when every row is copied
    void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t height) {
        if (width == dpitch && width == spitch) { width *= height; height = 1; dpitch = spitch = width; }
        const size_t total = width * height;
        const int parts = (total < ((size_t)1 << 20)) ? 1 : nthreads() + 1;
        if (parts == 1) { rows(dst, dpitch, src, spitch, width, 0, height, 0, width); return; }
        start();
        Batch b;
        b.left = parts - 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int i = 1; i < parts; ++i) jobs_.push_back(Job{dst, dpitch, src, spitch, width, height, i, parts, &b});
        }
        cv_.notify_all();
        run_part(Job{dst, dpitch, src, spitch, width, height, 0, parts, &b});
        std::unique_lock<std::mutex> lk(b.mu);
        b.cv.wait(lk, [&] { return b.left == 0; });
    }

  private:
    struct Batch { std::mutex mu; std::condition_variable cv; int left = 0; };
    struct Job { char* dst; size_t dpitch; const char* src; size_t spitch, width, height; int part, parts; Batch* batch; };
    static void rows(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t r0, size_t r1, size_t c0, size_t c1) {
        for (size_t r = r0; r < r1; ++r) std::memcpy(dst + r * dpitch + c0, src + r * spitch + c0, c1 - c0);
    }
    static void run_part(const Job& j) {
        if (j.height >= (size_t)j.parts) {      // split by rows
            const size_t r0 = j.height * (size_t)j.part / (size_t)j.parts, r1 = j.height * (size_t)(j.part + 1) / (size_t)j.parts;
            rows(j.dst, j.dpitch, j.src, j.spitch, j.width, r0, r1, 0, j.width);
        } else {                                // few long rows: split every row by columns (64-byte granules)
            size_t c0 = (j.width * (size_t)j.part / (size_t)j.parts) & ~(size_t)63, c1 = (j.width * (size_t)(j.part + 1) / (size_t)j.parts) & ~(size_t)63;
            if (j.part + 1 == j.parts) c1 = j.width;
            rows(j.dst, j.dpitch, j.src, j.spitch, j.width, 0, j.height, c0, c1);
        }
    }
    int nthreads() {
        static const int n = [] {
            unsigned hw = std::thread::hardware_concurrency();
            int t = (int)(hw ? hw : 8) / 2 - 1;
            return t < 1 ? 1 : (t > 7 ? 7 : t);
        }();
        return n;
    }
    void start() {
        std::lock_guard<std::mutex> lk(mu_);
        if (started_) return;
        started_ = true;
        for (int i = 0; i < nthreads(); ++i) {
            std::thread([this] {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk2(mu_);
                        cv_.wait(lk2, [&] { return !jobs_.empty(); });
                        j = jobs_.front();
                        jobs_.pop_front();
                    }
                    run_part(j);
                    { std::lock_guard<std::mutex> bl(j.batch->mu); --j.batch->left; }
                    j.batch->cv.notify_one();
                }
            }).detach();   // workers live as long as the process and sleep on the queue
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job> jobs_;
    bool started_ = false;
};
static CopyTeam& copy_team() {
    static CopyTeam* t = new CopyTeam();   // never destroyed: its detached workers may outlive static destruction
    return *t;
}
static bool host_pageable(const void* p) {
#ifdef AFB_HOST_EMU
    (void)p;
    return false;
#else
    // AFB_NO_RING=1: hand pageable pointers to the driver as they are (A/B runs of tools/time_host_paths.py)
    static const bool no_ring = [] { const char* e = std::getenv("AFB_NO_RING"); return e && e[0] == '1'; }();
    if (no_ring) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}

struct HostPipe {
#ifndef AFB_NPIPE
#define AFB_NPIPE 3          // slots of the pipeline (A/B: -DAFB_NPIPE=5 made no difference for the pageable ring)
#endif
    static constexpr int NPIPE = AFB_NPIPE;
    cudaStream_t stream[NPIPE] = {};
    void* buf[NPIPE] = {};
    size_t cap[NPIPE] = {};
    // pinned ring for pageable callers: one input and one output slot per stream
    void* hin[NPIPE] = {};
    void* hout[NPIPE] = {};
    size_t hin_cap[NPIPE] = {}, hout_cap[NPIPE] = {};
    int32_t ensure(int nbuf, size_t bytes) {
        for (int i = 0; i < nbuf; ++i) {
            if (!stream[i]) AFB_CUDA(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
            if (cap[i] < bytes) {
                if (buf[i]) { AFB_CUDA(cudaStreamSynchronize(stream[i])); cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
                AFB_CUDA(cudaMalloc(&buf[i], bytes));
                cap[i] = bytes;
            }
        }
        return AFB_OK;
    }
    int32_t ensure_pinned(int nbuf, size_t bytes, bool in, bool out) {
        for (int i = 0; i < nbuf; ++i) {
            if (in && hin_cap[i] < bytes) {
                if (hin[i]) cudaFreeHost(hin[i]);
                hin[i] = nullptr; hin_cap[i] = 0;
                AFB_CUDA(cudaHostAlloc(&hin[i], bytes, cudaHostAllocDefault));
                hin_cap[i] = bytes;
            }
            if (out && hout_cap[i] < bytes) {
                if (hout[i]) cudaFreeHost(hout[i]);
                hout[i] = nullptr; hout_cap[i] = 0;
                AFB_CUDA(cudaHostAlloc(&hout[i], bytes, cudaHostAllocDefault));
                hout_cap[i] = bytes;
            }
        }
        return AFB_OK;
    }
    void release() {
        for (int i = 0; i < NPIPE; ++i) {
            if (buf[i]) cudaFree(buf[i]);
            if (hin[i]) cudaFreeHost(hin[i]);
            if (hout[i]) cudaFreeHost(hout[i]);
            if (stream[i]) cudaStreamDestroy(stream[i]);
            buf[i] = nullptr; stream[i] = nullptr; cap[i] = 0;
            hin[i] = hout[i] = nullptr; hin_cap[i] = hout_cap[i] = 0;
        }
    }
};
struct PipePool {
    std::mutex mu;
    std::vector<HostPipe*> idle;
};
static PipePool g_pipes[MAX_DEVICES];
static constexpr size_t PIPES_KEPT = 4;  // idle pipes kept per device (each up to 3 x 64 MiB)
struct PipeLease {
    int dev;
    HostPipe* pipe = nullptr;
    PipeLease() : dev(current_device()) {
        std::lock_guard<std::mutex> lk(g_pipes[dev].mu);
        if (!g_pipes[dev].idle.empty()) { pipe = g_pipes[dev].idle.back(); g_pipes[dev].idle.pop_back(); }
        else pipe = new HostPipe();
    }
    ~PipeLease() {
        std::unique_lock<std::mutex> lk(g_pipes[dev].mu);
        if (g_pipes[dev].idle.size() < PIPES_KEPT) { g_pipes[dev].idle.push_back(pipe); return; }
        lk.unlock();
        pipe->release();
        delete pipe;
    }
    HostPipe* operator->() { return pipe; }
};
// ---- one large contiguous caller array to / from the device ---------------------------------------------------------
// Page-locked (or registered) memory is DMA-ed directly.  PAGEABLE memory would go through the driver's own bounce buffer on
// one thread (about 6-10 GB/s and synchronous); instead 32 MiB chunks go through the pinned ring of a leased pipe: the copy
// team fills / drains one slot while the DMA engine works on the neighbouring ones.
static constexpr size_t RING_CHUNK = (size_t)32 << 20;
static int32_t big_h2d(void* d, const void* h, size_t bytes) {
    if (bytes == 0) return AFB_OK;
    if (bytes < ((size_t)4 << 20) || !host_pageable(h)) { AFB_CUDA(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); return AFB_OK; }
    const int nbuf = (int)std::min<size_t>(HostPipe::NPIPE, (bytes + RING_CHUNK - 1) / RING_CHUNK);
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, 1));                    // streams
    AFB_TRY(pipe->ensure_pinned(nbuf, RING_CHUNK, true, false));
    int32_t st = AFB_OK;
    size_t c = 0;
    for (size_t off = 0; off < bytes && st == AFB_OK; off += RING_CHUNK, ++c) {
        const int si = (int)(c % nbuf);
        const size_t nb = std::min(RING_CHUNK, bytes - off);
        cudaError_t e = cudaStreamSynchronize(pipe->stream[si]);      // the slot's previous DMA has read it
        if (e != cudaSuccess) { st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); break; }
        copy_team().copy2d((char*)pipe->hin[si], nb, (const char*)h + off, nb, nb, 1);
        e = cudaMemcpyAsync((char*)d + off, pipe->hin[si], nb, cudaMemcpyHostToDevice, pipe->stream[si]);
        if (e != cudaSuccess) st = cuda_fail(e, "H2D", __FILE__, __LINE__);
    }
    for (int i = 0; i < nbuf; ++i) {
        cudaError_t e = cudaStreamSynchronize(pipe->stream[i]);
        if (e != cudaSuccess && st == AFB_OK) st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    return st;
}
// `after`: a stream whose work produced d (nullptr: the legacy stream, already ordered with the synchronous copies)
static int32_t big_d2h(void* h, const void* d, size_t bytes) {
    if (bytes == 0) return AFB_OK;
    if (bytes < ((size_t)4 << 20) || !host_pageable(h)) { AFB_CUDA(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); return AFB_OK; }
    AFB_CUDA(cudaDeviceSynchronize());                 // the pipe's streams are non-blocking: order them after the producer
    const size_t nchunks = (bytes + RING_CHUNK - 1) / RING_CHUNK;
    const int nbuf = (int)std::min<size_t>(HostPipe::NPIPE, nchunks);
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, 1));
    AFB_TRY(pipe->ensure_pinned(nbuf, RING_CHUNK, false, true));
    int32_t st = AFB_OK;
    auto issue = [&](size_t c) -> int32_t {
        const size_t off = c * RING_CHUNK, nb = std::min(RING_CHUNK, bytes - off);
        AFB_CUDA(cudaMemcpyAsync(pipe->hout[c % nbuf], (const char*)d + off, nb, cudaMemcpyDeviceToHost, pipe->stream[c % nbuf]));
        return AFB_OK;
    };
    for (size_t c = 0; c < (size_t)nbuf && st == AFB_OK; ++c) st = issue(c);
    for (size_t c = 0; c < nchunks && st == AFB_OK; ++c) {
        const int si = (int)(c % nbuf);
        const size_t off = c * RING_CHUNK, nb = std::min(RING_CHUNK, bytes - off);
        cudaError_t e = cudaStreamSynchronize(pipe->stream[si]);
        if (e != cudaSuccess) { st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); break; }
        copy_team().copy2d((char*)h + off, nb, (const char*)pipe->hout[si], nb, nb, 1);
        if (c + nbuf < nchunks) st = issue(c + nbuf);  // the slot is free again
    }
    for (int i = 0; i < nbuf; ++i) cudaStreamSynchronize(pipe->stream[i]);
    if (st != AFB_OK) cudaGetLastError();
    return st;
}

// pitched variant of big_d2h: `height` rows of `width` bytes, device pitch dpitch, host pitch hpitch
static int32_t big_d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t height) {
    if (width == 0 || height == 0) return AFB_OK;
    if (width * height < ((size_t)4 << 20) || !host_pageable(h)) {
        AFB_CUDA(cudaMemcpy2D(h, hpitch, d, dpitch, width, height, cudaMemcpyDeviceToHost));
        return AFB_OK;
    }
    AFB_CUDA(cudaDeviceSynchronize());
    const size_t rows_per = std::max<size_t>(1, RING_CHUNK / width);
    const size_t nchunks = (height + rows_per - 1) / rows_per;
    const int nbuf = (int)std::min<size_t>(HostPipe::NPIPE, nchunks);
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, 1));
    AFB_TRY(pipe->ensure_pinned(nbuf, rows_per * width, false, true));
    int32_t st = AFB_OK;
    auto issue = [&](size_t c) -> int32_t {
        const size_t r0 = c * rows_per, nr = std::min(rows_per, height - r0);
        AFB_CUDA(cudaMemcpy2DAsync(pipe->hout[c % nbuf], width, (const char*)d + r0 * dpitch, dpitch, width, nr,
                                   cudaMemcpyDeviceToHost, pipe->stream[c % nbuf]));
        return AFB_OK;
    };
    for (size_t c = 0; c < (size_t)nbuf && st == AFB_OK; ++c) st = issue(c);
    for (size_t c = 0; c < nchunks && st == AFB_OK; ++c) {
        const int si = (int)(c % nbuf);
        const size_t r0 = c * rows_per, nr = std::min(rows_per, height - r0);
        cudaError_t e = cudaStreamSynchronize(pipe->stream[si]);
        if (e != cudaSuccess) { st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); break; }
        copy_team().copy2d((char*)h + r0 * hpitch, hpitch, (const char*)pipe->hout[si], width, width, nr);
        if (c + nbuf < nchunks) st = issue(c + nbuf);
    }
    for (int i = 0; i < nbuf; ++i) cudaStreamSynchronize(pipe->stream[i]);
    if (st != AFB_OK) cudaGetLastError();
    return st;
}
// for the multi-device 2-D plan (afb_tensor.cu): one device thread each, current device = the slab's
namespace afb {
bool host_is_pageable(const void* p) { return host_pageable(p); }
int32_t host_big_h2d(void* d, const void* h, size_t bytes) { return big_h2d(d, h, bytes); }
int32_t host_big_d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t height) {
    return big_d2h_2d(h, hpitch, d, dpitch, width, height);
}
}  // namespace afb

static void pipes_release_all() {
    int keep = 0;
    const bool have = cudaGetDevice(&keep) == cudaSuccess;
    for (int d = 0; d < MAX_DEVICES; ++d) {
        std::lock_guard<std::mutex> lk(g_pipes[d].mu);
        if (g_pipes[d].idle.empty()) continue;
        cudaSetDevice(d);
        for (HostPipe* hp : g_pipes[d].idle) { hp->release(); delete hp; }
        g_pipes[d].idle.clear();
    }
    if (have) cudaSetDevice(keep);
    cudaGetLastError();
}

// ---- device block cache for the host-pointer plain functions (per device, grow-only up to a cap, freed by afb_shutdown) --
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> blocks;  // capacity -> block
    size_t bytes = 0;
};
static BlockCache g_cache[MAX_DEVICES];
static const size_t CACHE_MAX_BLOCK = (size_t)64 << 20;    // larger requests go straight to cudaMalloc / cudaFree
static const size_t CACHE_MAX_TOTAL = (size_t)512 << 20;
static int32_t block_cache_get(size_t bytes, void** out, size_t* cap) {
    if (bytes <= CACHE_MAX_BLOCK) {
        BlockCache& bc = g_cache[current_device()];
        std::lock_guard<std::mutex> lk(bc.mu);
        auto it = bc.blocks.lower_bound(bytes);
        if (it != bc.blocks.end() && it->first <= 4 * bytes + 4096) {
            *out = it->second; *cap = it->first;
            bc.bytes -= it->first;
            bc.blocks.erase(it);
            return AFB_OK;
        }
    }
    size_t want = bytes;
    if (bytes <= CACHE_MAX_BLOCK) { want = 4096; while (want < bytes) want <<= 1; }   // size classes: powers of two
    AFB_CUDA(cudaMalloc(out, want));
    *cap = want;
    return AFB_OK;
}
static void block_cache_put(void* p, size_t cap, int dev) {
    if (cap <= CACHE_MAX_BLOCK) {
        BlockCache& bc = g_cache[dev];
        std::lock_guard<std::mutex> lk(bc.mu);
        if (bc.bytes + cap <= CACHE_MAX_TOTAL) {
            bc.blocks.emplace(cap, p);
            bc.bytes += cap;
            return;
        }
    }
    cudaFree(p);
}
static void block_cache_release() {
    for (int d = 0; d < MAX_DEVICES; ++d) {
        std::lock_guard<std::mutex> lk(g_cache[d].mu);
        for (auto& kv : g_cache[d].blocks) cudaFree(kv.second);   // cudaFree works across devices of one process
        g_cache[d].blocks.clear();
        g_cache[d].bytes = 0;
    }
}

namespace afb {
// recycled device blocks for short-lived buffers (afb_buf_*): no cudaMalloc / cudaFree (and no device-wide sync) per temporary
int32_t device_block_get(size_t bytes, void** out, size_t* cap) { return block_cache_get(bytes ? bytes : 1, out, cap); }
void device_block_put(void* p, size_t cap, int dev) { block_cache_put(p, cap, dev); }
}  // namespace afb

// ---- fan-out over afb_init's device set -----------------------------------------------------------------
// Splits [0, total) into contiguous shares, one per device, and runs fn(share_index, lo, hi) on one host thread per device
// with that device current.  The first non-OK status (and its message) is returned to the caller's thread.
template <class Fn> static int32_t fanout(const std::vector<int>& devs, int64_t total, Fn fn) {
    const int G = (int)devs.size();
    std::vector<int32_t> st((size_t)G, AFB_OK);
    std::vector<std::string> msg((size_t)G);
    std::vector<std::thread> th;
    int keep = 0;
    AFB_CUDA(cudaGetDevice(&keep));
    for (int g = 0; g < G; ++g) {
        const int64_t q = total / G, r = total % G;
        const int64_t lo = g * q + std::min<int64_t>(g, r), hi = lo + q + (g < r ? 1 : 0);
        th.emplace_back([&, g, lo, hi]() {
            cudaError_t e = cudaSetDevice(devs[(size_t)g]);
            int32_t s = (e == cudaSuccess) ? ensure_device() : cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
            if (s == AFB_OK && hi > lo) s = fn(g, lo, hi);
            st[(size_t)g] = s;
            if (s != AFB_OK) { char buf[512]; afb_last_error(buf, sizeof buf); msg[(size_t)g] = buf; }
        });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(keep);
    for (int g = 0; g < G; ++g)
        if (st[(size_t)g] != AFB_OK) return fail(st[(size_t)g], msg[(size_t)g]);
    return AFB_OK;
}

// =====================================================================================================
// plans
// =====================================================================================================
enum PlanPath { PATH_EMPTY = 0, PATH_DIRECT = 1, PATH_LINE = 2, PATH_GENERIC = 3, PATH_TENSOR2D = 4, PATH_LARGE = 5, PATH_EXT = 6,
                PATH_IDENTITY = 7, PATH_PADUA = 8 };

struct afb_plan_s {
    int32_t kind = 0;
    int64_t n = 0, n2 = 0, batch = 0, stride = 0;
    uint32_t flags = 0;
    int path = PATH_EMPTY;
    int elem = 8;  // bytes per element (16 for Laurent)
    // LINE
    int Nc = 0;
    int line_kind = 0;
    const cplx* tw = nullptr;
    cplx* aux = nullptr;
    std::vector<cplx> K;  // MAXJ x MAXRL rotation constants (kernel parameters)
    double scale = 1.0;
    // DIRECT
    cplx* dtab = nullptr;
    // GENERIC (any n): length-n complex FFT engine + pre/post
    CfftPlan* cfft = nullptr;
    cplx* qtab = nullptr;      // exp(-i pi k/(2n)), k = 0..n  (Chebyshev)
    cplx* gwork = nullptr;     // gbatch x n complex
    int64_t gbatch = 0;
    int64_t cn = 0;            // EXT: complex length of the extension FFT
    // TENSOR2D
    afb_plan_s* col_plan = nullptr;  // length n,  batch n2
    afb_plan_s* row_plan = nullptr;  // length n2, batch n
    double* t2buf = nullptr;
    // PADUA: degree, index tables (grid position of every point; transposed-grid position of every coefficient),
    // the (d+2) x (d+1) grid W and its transpose T (= t2buf); col_plan / row_plan are second-kind (REDFT00) plans
    int64_t pdeg = 0;
    int32_t* pad_idx = nullptr;
    int32_t* tri_idx = nullptr;
    double* wbuf = nullptr;
    // scratch for in-place direct execution
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    std::mutex mu;
    int dev = afb::current_device();          // the device the plan's tables and work buffers live on
    // host-pointer execution over several devices: one replica plan per device of afb_init's set (made on first use)
    std::vector<afb_plan_s*> replicas;
    std::vector<int> replica_devs;
    void* mdev2d = nullptr;                   // in-process multi-device 2-D state (afb_tensor.cu)
};

// makes the plan's device current for the scope (a plan may be executed from any thread)
struct DeviceGuard {
    int keep = -1;
    explicit DeviceGuard(int dev) {
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) { keep = cur; cudaSetDevice(dev); }
    }
    ~DeviceGuard() { if (keep >= 0) cudaSetDevice(keep); }
};

static void plan_free(afb_plan_s* p) {
    if (!p) return;
    for (afb_plan_s* r : p->replicas) plan_free(r);
    if (p->mdev2d) mdev2d_destroy(p->mdev2d);
    DeviceGuard guard(p->dev);
    if (p->aux) cudaFree(p->aux);
    if (p->dtab) cudaFree(p->dtab);
    if (p->qtab) cudaFree(p->qtab);
    if (p->gwork) cudaFree(p->gwork);
    if (p->cfft) cfft_destroy(p->cfft);
    if (p->t2buf) cudaFree(p->t2buf);
    if (p->pad_idx) cudaFree(p->pad_idx);
    if (p->tri_idx) cudaFree(p->tri_idx);
    if (p->wbuf) cudaFree(p->wbuf);
    if (p->tmp) cudaFree(p->tmp);
    plan_free(p->col_plan);
    plan_free(p->row_plan);
    delete p;
}

static inline bool kind_is_real(int kind) { return kind <= AFB_FOURIER_INV || kind >= AFB_CHEB1_2D_FWD; }

// Per-plan twiddle tables and rotation constants of the LINE path (see dit_last_paired in afb_transform.cu).
// Every twiddle family is t_j(k) = g_j * exp(-i pi m_j k / (2n)):
//   aux[J*pi + j] = t_j(pi) for pi = 0..NSL/2, aux[J*(NSL/2+1) + j] = t_j(N/2)
//   K[j][q] = exp(-i pi m_j NSL q/(2n))                       (2q <  RL: t_j(pi + NSL q)   = t_j(pi) K)
//           = (g_j/conj g_j) exp(-i pi m_j (N - NSL q)/(2n))  (2q >= RL: t_j(N-pi-NSL q) = conj(t_j(pi)) K)
struct TwFamily { long double gre, gim; int64_t m; };

static cplx tw_eval(const TwFamily& f, int64_t k, int64_t n, long double sre = 1.0L, long double sim = 0.0L) {
    const long double PI = 3.141592653589793238462643383279502884L;
    int64_t e = (int64_t)(((__int128)f.m * k) % (4 * n));
    if (e < 0) e += 4 * n;
    const long double a = PI * (long double)e / (2.0L * (long double)n);
    const long double wr = cosl(a), wi = -sinl(a);
    const long double gr = sre, gi = sim;  // prefactor
    return make_double2((double)(gr * wr - gi * wi), (double)(gr * wi + gi * wr));
}

static int32_t build_line_aux(afb_plan_s* p) {
    const int64_t n = p->n;
    const int64_t N = p->Nc;
    const long double inv_n = 1.0L / (long double)n;
    TwFamily fam[2];
    int J = 0;
    switch (p->kind) {
        case AFB_CHEB1_FWD:
            fam[0] = {inv_n, 0.0L, 1}; fam[1] = {0.0L, -inv_n, 5}; J = 2;
            p->line_kind = ILK_CHEB_FWD; p->scale = 1.0 / (double)n; break;
        case AFB_CHEB1_INV:
            fam[0] = {1.0L, 0.0L, -1}; fam[1] = {1.0L, 0.0L, -4}; J = 2;
            p->line_kind = ILK_CHEB_INV; break;
        case AFB_FOURIER_FWD:
            fam[0] = {0.0L, -inv_n, 4}; J = 1;
            p->line_kind = ILK_FOURIER_FWD; p->scale = 1.0 / (double)n; break;
        case AFB_FOURIER_INV:
            fam[0] = {1.0L, 0.0L, -4}; J = 1;
            p->line_kind = ILK_FOURIER_INV; break;
        case AFB_LAURENT_FWD:
            p->line_kind = ILK_LAURENT_FWD; p->scale = 1.0 / (double)n; return AFB_OK;
        case AFB_LAURENT_INV:
            p->line_kind = ILK_LAURENT_INV; return AFB_OK;
        default:
            return fail(AFB_BAD_ARG, "bad kind for line path");
    }
    int RL = 0, NSL = 0;
    AFB_TRY(line_last_pass((int)N, &RL, &NSL));
    std::vector<cplx> h((size_t)J * (size_t)(NSL / 2 + 2));
    for (int j = 0; j < J; ++j) {
        for (int64_t pi = 0; pi <= NSL / 2; ++pi) h[(size_t)(J * pi + j)] = tw_eval(fam[j], pi, n, fam[j].gre, fam[j].gim);
        h[(size_t)(J * (NSL / 2 + 1) + j)] = tw_eval(fam[j], N / 2, n, fam[j].gre, fam[j].gim);
    }
    p->K.assign(2 * LINE_MAXRL, make_double2(1.0, 0.0));
    for (int j = 0; j < J; ++j) {
        const bool g_imag = fam[j].gim != 0.0L;  // g / conj(g) = -1 for purely imaginary g, +1 for real g
        for (int q = 0; q < RL; ++q) {
            if (2 * q < RL) p->K[(size_t)(j * LINE_MAXRL + q)] = tw_eval(fam[j], (int64_t)NSL * q, n);
            else p->K[(size_t)(j * LINE_MAXRL + q)] = tw_eval(fam[j], N - (int64_t)NSL * q, n, g_imag ? -1.0L : 1.0L, 0.0L);
        }
    }
    return upload_table(h, &p->aux);
}

static int32_t build_direct_tab(afb_plan_s* p) {
    const int64_t n = p->n;
    std::vector<cplx> h;
    const long double PI = 3.141592653589793238462643383279502884L;
    if (p->kind == AFB_CHEB1_FWD || p->kind == AFB_CHEB1_INV) {
        h.resize((size_t)(4 * n));
        for (int64_t m = 0; m < 4 * n; ++m) h[m] = make_double2((double)cosl(PI * m / (2.0L * n)), 0.0);
    } else {
        h.resize((size_t)n);
        for (int64_t m = 0; m < n; ++m) {
            const long double a = 2.0L * PI * m / (long double)n;
            h[m] = make_double2((double)cosl(a), (double)(-sinl(a)));
        }
    }
    return upload_table(h, &p->dtab);
}

static int32_t build_generic(afb_plan_s* p) {
    const int64_t n = p->n;
    // sub-batch: keep the complex work buffer below ~256 MiB
    int64_t gb = std::max<int64_t>(1, (int64_t)(256ll << 20) / (16 * n));
    gb = std::min<int64_t>(gb, std::max<int64_t>(p->batch, 1));
    p->gbatch = gb;
    AFB_TRY(cfft_create(&p->cfft, n, gb));
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, sizeof(cplx) * (size_t)(gb * n)));
    p->gwork = reinterpret_cast<cplx*>(d);
    if (p->kind == AFB_CHEB1_FWD || p->kind == AFB_CHEB1_INV) {
        std::vector<cplx> h((size_t)n + 1);
        const long double PI = 3.141592653589793238462643383279502884L;
        for (int64_t k = 0; k <= n; ++k) {
            const long double a = PI * k / (2.0L * (long double)n);
            h[k] = make_double2((double)cosl(a), (double)(-sinl(a)));
        }
        AFB_TRY(upload_table(h, &p->qtab));
    }
    return AFB_OK;
}

static int32_t plan_create_1d(afb_plan_s** out, int32_t kind, int64_t n, int64_t batch, int64_t stride, uint32_t flags) {
    afb_plan_s* p = new afb_plan_s();
    p->kind = kind; p->n = n; p->batch = batch; p->stride = stride; p->flags = flags;
    p->elem = (kind == AFB_LAURENT_FWD || kind == AFB_LAURENT_INV || kind >= AFB_TAYLOR_FWD) ? 16 : 8;
    int32_t st = AFB_OK;
    if (n == 0 || batch == 0) {
        p->path = PATH_EMPTY;
    } else if (kind >= AFB_TAYLOR_FWD) {
        p->path = PATH_GENERIC;   // plain complex FFT with a copy / reversal on either side
        st = build_generic(p);
    } else if (kind >= AFB_CHEB2_FWD) {
        // extension kinds: REDFT00 through an even extension of length 2(n-1), SinSpace through an odd one of 2n+2
        const bool cheb2 = kind == AFB_CHEB2_FWD || kind == AFB_CHEB2_INV;
        if (cheb2 && n == 1) {
            p->path = PATH_IDENTITY;
        } else {
            p->path = PATH_EXT;
            p->cn = cheb2 ? 2 * (n - 1) : 2 * n + 2;
            int64_t gb = std::max<int64_t>(1, (int64_t)(256ll << 20) / (16 * p->cn));
            gb = std::min<int64_t>(gb, batch);
            p->gbatch = gb;
            st = cfft_create(&p->cfft, p->cn, gb);
            if (st == AFB_OK) {
                void* d = nullptr;
                cudaError_t e = cudaMalloc(&d, sizeof(cplx) * (size_t)(gb * p->cn));
                if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(ext work)", __FILE__, __LINE__);
                p->gwork = reinterpret_cast<cplx*>(d);
            }
        }
    } else if (n <= DIRECT_MAX) {
        p->path = PATH_DIRECT;
        st = build_direct_tab(p);
    } else {
        const int64_t Nc = (p->elem == 16) ? n : n / 2;
        if (is_pow2(n) && Nc >= LINE_MIN_N && Nc <= LINE_MAX_N) {
            p->path = PATH_LINE;
            p->Nc = (int)Nc;
            st = root_table(Nc, &p->tw);
            if (st == AFB_OK) st = build_line_aux(p);
        } else if (is_pow2(n) && p->elem == 8 && Nc > LINE_MAX_N) {
            // large power-of-two real transform: half-length four-step FFT + one pair kernel
            p->path = PATH_LARGE;
            p->Nc = (int)Nc;
            st = cfft_create(&p->cfft, Nc, 1);
            if (st == AFB_OK) {
                void* d = nullptr;
                cudaError_t e = cudaMalloc(&d, sizeof(cplx) * (size_t)Nc);
                if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(large work)", __FILE__, __LINE__);
                p->gwork = reinterpret_cast<cplx*>(d);
            }
        } else {
            p->path = PATH_GENERIC;
            st = build_generic(p);
        }
    }
    if (st != AFB_OK) { plan_free(p); return st; }
    *out = p;
    return AFB_OK;
}

static int32_t exec_1d_device(afb_plan_s* p, const void* d_in, void* d_out, int64_t batch, int64_t istride,
                              int64_t ostride, void* stream, int64_t oes = 1) {
    switch (p->path) {
        case PATH_EMPTY: return AFB_OK;
        case PATH_DIRECT: {
            const void* src = d_in;
            if (d_in == d_out) {  // the O(n^2) kernel is not in-place safe: stage the input
                const size_t bytes = (size_t)p->elem * (size_t)((batch - 1) * istride + p->n);
                if (bytes > p->tmp_bytes) {
                    AFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
                    if (p->tmp) cudaFree(p->tmp);
                    p->tmp = nullptr; p->tmp_bytes = 0;
                    AFB_CUDA(cudaMalloc(&p->tmp, bytes));
                    p->tmp_bytes = bytes;
                }
                AFB_CUDA(cudaMemcpyAsync(p->tmp, d_in, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
                src = p->tmp;
            }
            return launch_direct(p->kind, (int)p->n, src, d_out, batch, istride, ostride, oes, p->dtab, stream);
        }
        case PATH_LINE:
            return launch_line(p->line_kind, p->Nc, d_in, d_out, batch, istride, ostride, oes, p->tw, p->aux,
                               p->K.empty() ? nullptr : p->K.data(), p->scale, stream);
        case PATH_LARGE: {
            if (oes != 1) return fail(AFB_UNSUPPORTED, "large path has no strided output");
            const bool cheb = p->kind == AFB_CHEB1_FWD || p->kind == AFB_CHEB1_INV;
            for (int64_t b = 0; b < batch; ++b) {
                const double* in = reinterpret_cast<const double*>(d_in) + b * istride;
                double* out = reinterpret_cast<double*>(d_out) + b * ostride;
                if ((p->kind & 1) == 0) {
                    AFB_TRY(cfft_exec_dctorder(p->cfft, in, p->gwork, false, false, 1.0, cheb ? 1 : 0, stream));
                    AFB_TRY(launch_large_post(p->kind, p->n, p->gwork, out, stream));
                } else {
                    AFB_TRY(launch_large_pre(p->kind, p->n, in, p->gwork, stream));
                    AFB_TRY(cfft_exec_dctorder(p->cfft, p->gwork, out, false, true, 1.0, cheb ? 2 : 0, stream));
                }
            }
            return AFB_OK;
        }
        case PATH_IDENTITY:
            AFB_CUDA(cudaMemcpy2DAsync(d_out, 8 * (size_t)ostride, d_in, 8 * (size_t)istride, 8, (size_t)batch,
                                       cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            return AFB_OK;
        case PATH_EXT: {
            if (oes != 1) return fail(AFB_UNSUPPORTED, "extension path has no strided output");
            if (cfft_ext_fused_ok(p->cfft))
                return launch_ext_fused(p->cfft, p->kind, p->n, reinterpret_cast<const double*>(d_in), istride,
                                        reinterpret_cast<double*>(d_out), ostride, batch, stream);
            for (int64_t b0 = 0; b0 < batch; b0 += p->gbatch) {
                const int64_t nb = std::min(p->gbatch, batch - b0);
                const double* in = reinterpret_cast<const double*>(d_in) + b0 * istride;
                double* out = reinterpret_cast<double*>(d_out) + b0 * ostride;
                AFB_TRY(launch_ext_pre(p->kind, p->n, p->cn, in, istride, p->gwork, nb, stream, p->cfft));
                AFB_TRY(cfft_exec_work(p->cfft, p->gwork, (nb + 1) / 2, false, stream));  // two lines per FFT
                AFB_TRY(launch_ext_post(p->kind, p->n, p->cn, p->gwork, out, ostride, nb, stream, p->cfft));
            }
            return AFB_OK;
        }
        case PATH_GENERIC: {
            if (oes != 1) return fail(AFB_UNSUPPORTED, "generic path has no strided output");
            const bool inv = (p->kind & 1) != 0;
            if (p->kind <= AFB_FOURIER_INV && cfft_real_fused_ok(p->cfft))
                return launch_blue_real(p->cfft, p->kind, reinterpret_cast<const double*>(d_in), istride,
                                        reinterpret_cast<double*>(d_out), ostride, batch, p->qtab, stream);
            for (int64_t b0 = 0; b0 < batch; b0 += p->gbatch) {
                const int64_t nb = std::min(p->gbatch, batch - b0);
                const char* in = reinterpret_cast<const char*>(d_in) + (size_t)p->elem * (size_t)(b0 * istride);
                char* out = reinterpret_cast<char*>(d_out) + (size_t)p->elem * (size_t)(b0 * ostride);
                AFB_TRY(launch_generic_pre(p->kind, p->n, in, istride, p->gwork, nb, p->qtab, stream, p->cfft));
                // the four real kinds ride two lines per complex FFT (see generic_pre_kernel)
                const int64_t nfft = (p->kind <= AFB_FOURIER_INV) ? (nb + 1) / 2 : nb;
                AFB_TRY(cfft_exec_work(p->cfft, p->gwork, nfft, inv, stream));
                AFB_TRY(launch_generic_post(p->kind, p->n, p->gwork, out, ostride, nb, p->qtab, stream, p->cfft));
            }
            return AFB_OK;
        }
        default: return fail(AFB_BAD_ARG, "plan path");
    }
}

static int32_t plan_launch_count(afb_plan_s* p) {
    switch (p->path) {
        case PATH_EMPTY: return 0;
        case PATH_DIRECT: return 1;
        case PATH_LINE: return 1;
        case PATH_LARGE: return (int32_t)(3 * p->batch);
        case PATH_IDENTITY: return 0;
        case PATH_PADUA: return 4 + plan_launch_count(p->col_plan) + plan_launch_count(p->row_plan);  // memset, scatter, transpose, gather
        case PATH_EXT: {
            if (cfft_ext_fused_ok(p->cfft)) return 1;
            const int64_t reps = (p->batch + p->gbatch - 1) / p->gbatch;
            return (int32_t)(reps * (2 + cfft_work_launches(p->cfft)));
        }
        case PATH_GENERIC: {
            if (p->kind <= AFB_FOURIER_INV && cfft_real_fused_ok(p->cfft)) return 1;
            const int64_t reps = (p->batch + p->gbatch - 1) / p->gbatch;
            return (int32_t)(reps * (2 + cfft_work_launches(p->cfft)));
        }
        case PATH_TENSOR2D: {
            if (!p->row_plan) return plan_launch_count(p->col_plan) + 2;  // + outer and inner row kernels
            return plan_launch_count(p->col_plan) + plan_launch_count(p->row_plan) + 2;
        }
    }
    return 0;
}

// ---- Padua plans ------------------------------------------------------------------
// Degree d with (d+1)(d+2)/2 >= N  (upstream: Int(cld(-3 + sqrt(1 + 8N), 2))).
static int64_t padua_degree(int64_t N) {
    int64_t d = (int64_t)std::ceil((-3.0 + std::sqrt(1.0 + 8.0 * (double)N)) / 2.0);
    if (d < 0) d = 0;
    while ((d + 1) * (d + 2) / 2 < N) ++d;
    while (d > 0 && d * (d + 1) / 2 >= N) --d;
    return d;
}
// Grid position iy + (d+2) ix of every Padua point, in the emission order of upstream `paduapoints`
// (k = d..0: x = -cospi(k/d) = cos(ix pi/d), ix = d-k;  j = NN+delta..1: y = cos(iy pi/(d+1)) with
//  iy = d+2-2j for odd d-k, d+3-2j for even d-k).
static std::vector<int32_t> padua_grid_index(int64_t d) {
    std::vector<int32_t> idx;
    idx.reserve((size_t)((d + 1) * (d + 2) / 2));
    const int64_t NN = d / 2 + 1;
    for (int64_t k = d; k >= 0; --k) {
        const int64_t delta = (d % 2 == 1) ? (k % 2) : 0;
        for (int64_t j = NN + delta; j >= 1; --j) {
            const int64_t ix = d - k, iy = ((d - k) % 2 == 1) ? d + 2 - 2 * j : d + 3 - 2 * j;
            idx.push_back((int32_t)(iy + (d + 2) * ix));
        }
    }
    return idx;
}
static int32_t upload_idx(const std::vector<int32_t>& h, int32_t** dev) {
    void* dptr = nullptr;
    AFB_CUDA(cudaMalloc(&dptr, sizeof(int32_t) * std::max<size_t>(h.size(), 1)));
    cudaError_t e = cudaMemcpy(dptr, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dptr); return cuda_fail(e, "cudaMemcpy(index table)", __FILE__, __LINE__); }
    *dev = reinterpret_cast<int32_t*>(dptr);
    return AFB_OK;
}
static int32_t plan_create_padua(afb_plan_s** out, int32_t kind, int64_t N, uint32_t flags) {
    const int64_t d = padua_degree(N);
    if ((d + 1) * (d + 2) / 2 != N)
        return fail(AFB_BAD_SIZE, "Padua transforms can only be applied to vectors of length (n+1)*(n+2)/2");
    if ((d + 2) * (d + 1) >= ((int64_t)1 << 31)) return fail(AFB_BAD_SIZE, "Padua transform: degree too large");
    afb_plan_s* p = new afb_plan_s();
    p->kind = kind; p->n = N; p->batch = 1; p->stride = N; p->flags = flags; p->pdeg = d;
    p->path = (N == 0) ? PATH_EMPTY : (d == 0 ? PATH_IDENTITY : PATH_PADUA);
    int32_t st = AFB_OK;
    if (p->path == PATH_PADUA) {
        const int64_t m = d + 2, l = d + 1;  // grid: m y-points (rows, contiguous) x l x-points
        const int32_t k1 = (kind == AFB_PADUA_FWD) ? AFB_CHEB2_FWD : AFB_CHEB2_INV;
        st = plan_create_1d(&p->col_plan, k1, m, l, m, flags);               // along y on W (m x l)
        if (st == AFB_OK) st = plan_create_1d(&p->row_plan, k1, l, m, l, flags);  // along x on T = W^T (l x m)
        if (st == AFB_OK) st = upload_idx(padua_grid_index(d), &p->pad_idx);
        if (st == AFB_OK) {
            // coefficient l of the total-degree order (diagonal s, x-degree i = 0..s, y-degree s - i) sits at T[i + l*j]
            std::vector<int32_t> tri;
            tri.reserve((size_t)N);
            for (int64_t sdeg = 0; sdeg <= d; ++sdeg)
                for (int64_t i = 0; i <= sdeg; ++i) tri.push_back((int32_t)(i + l * (sdeg - i)));
            st = upload_idx(tri, &p->tri_idx);
        }
        for (double** buf : {&p->wbuf, &p->t2buf}) {
            if (st != AFB_OK) break;
            void* dptr = nullptr;
            cudaError_t e = cudaMalloc(&dptr, sizeof(double) * (size_t)(m * l));
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(Padua grid)", __FILE__, __LINE__);
            *buf = reinterpret_cast<double*>(dptr);
        }
    }
    if (st != AFB_OK) { plan_free(p); return st; }
    *out = p;
    return AFB_OK;
}
// forward:  zero grid <- values; REDFT00 along y; transpose; REDFT00 along x; coefficients = 2 * triangle
// inverse:  zero grid^T <- coefficients; evaluate along x; transpose; evaluate along y; values = grid[points]
static int32_t exec_padua_device(afb_plan_s* p, const double* d_in, double* d_out, void* stream) {
    const int64_t d = p->pdeg, m = d + 2, l = d + 1, N = p->n;
    cudaStream_t s = (cudaStream_t)stream;
    if (p->kind == AFB_PADUA_FWD) {
        AFB_CUDA(cudaMemsetAsync(p->wbuf, 0, sizeof(double) * (size_t)(m * l), s));
        AFB_TRY(launch_scatter_idx(p->wbuf, d_in, p->pad_idx, N, stream));
        AFB_TRY(exec_1d_device(p->col_plan, p->wbuf, p->wbuf, l, m, m, stream, 1));
        AFB_TRY(launch_transpose(p->wbuf, p->t2buf, m, l, stream));
        AFB_TRY(exec_1d_device(p->row_plan, p->t2buf, p->t2buf, m, l, l, stream, 1));
        return launch_gather_idx(d_out, p->t2buf, p->tri_idx, N, 2.0, stream);
    }
    AFB_CUDA(cudaMemsetAsync(p->t2buf, 0, sizeof(double) * (size_t)(m * l), s));
    AFB_TRY(launch_scatter_idx(p->t2buf, d_in, p->tri_idx, N, stream));
    AFB_TRY(exec_1d_device(p->row_plan, p->t2buf, p->t2buf, m, l, l, stream, 1));
    AFB_TRY(launch_transpose(p->t2buf, p->wbuf, l, m, stream));
    AFB_TRY(exec_1d_device(p->col_plan, p->wbuf, p->wbuf, l, m, m, stream, 1));
    return launch_gather_idx(d_out, p->wbuf, p->pad_idx, N, 1.0, stream);
}

namespace afb {
int32_t slab_col_chunk(afb_plan p, const double* d_in, double* d_out, int64_t batch, void* stream) {
    if (!p || batch < 0 || batch > p->batch) return fail(AFB_BAD_ARG, "slab_col_chunk: bad argument");
    DeviceGuard guard(p->dev);
    return exec_1d_device(p, d_in, d_out, batch, p->stride, p->stride, stream);
}
}  // namespace afb

extern "C" {

int32_t afb_version(void) { return AFB_VERSION; }

int32_t afb_last_error(char* buf, size_t len) {
    if (buf && len) {
        std::strncpy(buf, g_err_msg.c_str(), len - 1);
        buf[len - 1] = 0;
    }
    return g_err_code;
}

int32_t afb_init(const int32_t* devices, int32_t ndev) {
    if (ndev < 0 || (ndev > 0 && !devices)) return fail(AFB_BAD_ARG, "afb_init: bad device list");
    if (ndev > MAX_DEVICES) return fail(AFB_UNSUPPORTED, "afb_init: at most 16 devices");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(AFB_NO_DEVICE, "no CUDA device visible: libapproxfun_b200 has no CPU fallback");
    }
    std::vector<int> list;
    for (int32_t i = 0; i < ndev; ++i) {
        if (devices[i] < 0 || devices[i] >= n || devices[i] >= MAX_DEVICES) return fail(AFB_BAD_ARG, "afb_init: device index out of range");
        for (int d : list) if (d == devices[i]) return fail(AFB_BAD_ARG, "afb_init: device listed twice");
        list.push_back(devices[i]);
    }
    if (!list.empty()) {
#ifndef AFB_HOST_EMU
        // every pair of the set talks over NVLink peer memory (the 2-D exchange writes straight into the peers' buffers)
        for (int a : list) {
            AFB_CUDA(cudaSetDevice(a));
            AFB_TRY(ensure_device());
            for (int b : list) {
                if (a == b) continue;
                int can = 0;
                AFB_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
                if (!can) continue;
                cudaError_t pe = cudaDeviceEnablePeerAccess(b, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(pe, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
        }
#endif
        AFB_CUDA(cudaSetDevice(list[0]));
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_devlist = list.size() > 1 ? list : std::vector<int>();
    }
    return ensure_device();
}

// Drops the staging pipes and the block cache of every device.  Plans stay valid (the twiddle tables they point into live
// as long as the process); the device set of afb_init is kept.
int32_t afb_shutdown(void) {
    pipes_release_all();
    block_cache_release();
    return AFB_OK;
}

int32_t afb_device_info(int32_t* sms, int64_t* hbm_bytes, char* name, size_t name_len) {
    AFB_TRY(ensure_device());
    int dev = 0;
    AFB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    AFB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    if (name && name_len) { std::strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return AFB_OK;
}

// ---- memory helpers ---------------------------------------------------------------------------------
int32_t afb_malloc(void** dptr, size_t bytes) {
    if (!dptr) return fail(AFB_BAD_ARG, "afb_malloc: null out pointer");
    AFB_TRY(ensure_device());
    AFB_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return AFB_OK;
}
int32_t afb_free(void* dptr) { if (dptr) AFB_CUDA(cudaFree(dptr)); return AFB_OK; }
int32_t afb_host_alloc(void** hptr, size_t bytes) {
    if (!hptr) return fail(AFB_BAD_ARG, "afb_host_alloc: null out pointer");
    AFB_TRY(ensure_device());
    AFB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return AFB_OK;
}
int32_t afb_host_free(void* hptr) { if (hptr) AFB_CUDA(cudaFreeHost(hptr)); return AFB_OK; }
int32_t afb_host_register(void* hptr, size_t bytes) {
    if (!hptr || bytes == 0) return fail(AFB_BAD_ARG, "afb_host_register: null pointer or zero size");
    AFB_TRY(ensure_device());
    AFB_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterPortable));
    return AFB_OK;
}
int32_t afb_host_unregister(void* hptr) {
    if (!hptr) return fail(AFB_BAD_ARG, "afb_host_unregister: null pointer");
    AFB_TRY(ensure_device());
    AFB_CUDA(cudaHostUnregister(hptr));
    return AFB_OK;
}
int32_t afb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    AFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return AFB_OK;
}
int32_t afb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    AFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return AFB_OK;
}
int32_t afb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    AFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return AFB_OK;
}
int32_t afb_stream_sync(void* stream) { AFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return AFB_OK; }

// ---- plans --------------------------------------------------------------------------------------------
int32_t afb_plan_create(afb_plan* out, int32_t kind, int64_t n, int64_t n2, int64_t batch, int64_t stride,
                        uint32_t flags) {
    if (!out) return fail(AFB_BAD_ARG, "afb_plan_create: null out pointer");
    *out = nullptr;
    if (kind < AFB_CHEB1_FWD || kind > AFB_LAURENT_CHEB1_2D_INV) return fail(AFB_BAD_ARG, "afb_plan_create: unknown kind");
    if (n < 0 || n2 < 0 || batch < 0) return fail(AFB_BAD_SIZE, "afb_plan_create: negative size");
    AFB_TRY(ensure_device());
    if (kind == AFB_CHEB1_2D_FWD || kind == AFB_CHEB1_2D_INV || kind >= AFB_FOURIER_CHEB1_2D_FWD) {
        if (batch != 1) return fail(AFB_BAD_SIZE, "afb_plan_create: 2D plans take batch == 1");
        // factor-1 plan on the columns, Chebyshev plan on the rows.  A complex (Laurent) matrix is a real 2n x n2 matrix whose
        // rows are the real and imaginary parts of the complex rows, and the row transform is real-linear.
        const bool inv = (kind & 1) != 0;
        const int32_t kcol = (kind <= AFB_CHEB1_2D_INV) ? (inv ? AFB_CHEB1_INV : AFB_CHEB1_FWD)
                           : (kind <= AFB_FOURIER_CHEB1_2D_INV) ? (inv ? AFB_FOURIER_INV : AFB_FOURIER_FWD)
                                                                : (inv ? AFB_LAURENT_INV : AFB_LAURENT_FWD);
        const int32_t krow = inv ? AFB_CHEB1_INV : AFB_CHEB1_FWD;
        afb_plan_s* p = new afb_plan_s();
        p->kind = kind; p->n = n; p->n2 = n2; p->batch = 1; p->stride = n; p->flags = flags;
        p->elem = (kcol == AFB_LAURENT_FWD || kcol == AFB_LAURENT_INV) ? 16 : 8;
        p->path = (n == 0 || n2 == 0) ? PATH_EMPTY : PATH_TENSOR2D;
        if (p->path == PATH_TENSOR2D) {
            const int64_t nreal = n * (p->elem / 8);
            int32_t st = plan_create_1d(&p->col_plan, kcol, n, n2, n, flags);
            // row pass: strided four-step on row pairs when the shape allows, else transpose + lines + transpose
            if (st == AFB_OK && ((flags & AFB_PLAN_2D_TRANSPOSE) || !rows_cheb_supported(nreal, n2)))
                st = plan_create_1d(&p->row_plan, krow, n2, nreal, n2, flags);
            if (st == AFB_OK) {
                void* d = nullptr;
                const int64_t tdoubles = p->row_plan ? nreal * n2 : rows_tmp_doubles(nreal, n2);
                cudaError_t e = cudaMalloc(&d, sizeof(double) * (size_t)tdoubles);
                if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(2D transpose buffer)", __FILE__, __LINE__);
                p->t2buf = reinterpret_cast<double*>(d);
            }
            if (st != AFB_OK) { plan_free(p); return st; }
        }
        *out = p;
        return AFB_OK;
    }
    if (n2 != 0) return fail(AFB_BAD_SIZE, "afb_plan_create: n2 must be 0 for 1D kinds");
    if (kind == AFB_PADUA_FWD || kind == AFB_PADUA_INV) {
        if (batch != 1) return fail(AFB_BAD_SIZE, "afb_plan_create: Padua plans take batch == 1");
        return plan_create_padua(out, kind, n, flags);
    }
    if (stride < n) return fail(AFB_BAD_SIZE, "afb_plan_create: stride < n");
    if ((kind == AFB_CHEB1_FWD || kind == AFB_CHEB1_INV || kind == AFB_FOURIER_FWD || kind == AFB_FOURIER_INV) &&
        (stride & 1) && batch > 1 && n > DIRECT_MAX && is_pow2(n))
        return fail(AFB_BAD_SIZE, "afb_plan_create: odd stride breaks 16-byte alignment of the vectorised path");
    return plan_create_1d(out, kind, n, batch, stride, flags);
}

int32_t afb_plan_destroy(afb_plan plan) {
    plan_free(plan);
    return AFB_OK;
}

int32_t afb_plan_io_doubles(afb_plan p, int64_t* in_doubles, int64_t* out_doubles) {
    if (!p || !in_doubles || !out_doubles) return fail(AFB_BAD_ARG, "afb_plan_io_doubles: null argument");
    int64_t elems;
    if (p->path == PATH_EMPTY) elems = 0;
    else if (p->path == PATH_TENSOR2D) elems = p->n * p->n2;
    else if (p->path == PATH_PADUA) elems = p->n;
    else elems = (p->batch - 1) * p->stride + p->n;
    *in_doubles = *out_doubles = elems * (p->elem / 8);
    return AFB_OK;
}

int32_t afb_plan_launches(afb_plan plan, int32_t* launches) {
    if (!plan || !launches) return fail(AFB_BAD_ARG, "afb_plan_launches: null argument");
    *launches = plan_launch_count(plan);
    return AFB_OK;
}

int32_t afb_plan_execute_device(afb_plan p, const void* d_in, void* d_out, void* stream) {
    if (!p) return fail(AFB_BAD_ARG, "afb_plan_execute_device: null plan");
    if (p->path == PATH_EMPTY) return AFB_OK;
    if (!d_in || !d_out) return fail(AFB_BAD_ARG, "afb_plan_execute_device: null buffer");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(AFB_BAD_ARG, "afb_plan_execute_device: buffers must be 16-byte aligned");
    std::lock_guard<std::mutex> lk(p->mu);
    DeviceGuard guard(p->dev);
    if (p->path == PATH_TENSOR2D) {
        // C = T_x V T_y^T on a column-major n x n2 matrix.
        const int64_t n = p->n, n2 = p->n2, nreal = n * (p->elem / 8);
        const bool inv = (p->kind & 1) != 0;
        // Column pass on contiguous lines, then the row pass.  (Fusing the transposes into the line kernels' stores --
        // 8-byte scatters with a 128 KB stride -- was measured 1.5x SLOWER at 16384^2: 6.06 ms vs 3.97 ms.)
        AFB_TRY(exec_1d_device(p->col_plan, d_in, d_out, n2, n, n, stream));
        if (!p->row_plan)
            return launch_rows_cheb(inv, reinterpret_cast<const double*>(d_out), p->t2buf, reinterpret_cast<double*>(d_out),
                                    nreal, n2, stream);
        AFB_TRY(launch_transpose(reinterpret_cast<const double*>(d_out), p->t2buf, nreal, n2, stream));
        AFB_TRY(exec_1d_device(p->row_plan, p->t2buf, p->t2buf, nreal, n2, n2, stream));
        AFB_TRY(launch_transpose(p->t2buf, reinterpret_cast<double*>(d_out), n2, nreal, stream));
        return AFB_OK;
    }
    if (p->path == PATH_PADUA)
        return exec_padua_device(p, reinterpret_cast<const double*>(d_in), reinterpret_cast<double*>(d_out), stream);
    return exec_1d_device(p, d_in, d_out, p->batch, p->stride, p->stride, stream);
}

}  // extern "C"

// One device: chunked three-stage pipeline H2D | kernels | D2H on up to NPIPE streams of a leased pipe, packed
// (stride = n) on the device.  `batch` lines starting at in / out (the plan's stride apart).
static int32_t host_pipeline_1d(afb_plan_s* p, const void* in, void* out, int64_t batch) {
    const int64_t n = p->n;
    const size_t line_bytes = (size_t)p->elem * (size_t)n;
    const size_t pitch = (size_t)p->elem * (size_t)p->stride;
    // pageable caller arrays of some size go through the pinned ring (see CopyTeam); pinned / registered ones are DMA-ed directly
    const bool big = line_bytes * (size_t)batch >= ((size_t)4 << 20);
    const bool bounce_in = big && host_pageable(in), bounce_out = big && host_pageable(out);
    const int64_t chunk_target = (bounce_in || bounce_out) ? ((int64_t)32 << 20) : ((int64_t)64 << 20);
    int64_t ch = std::max<int64_t>(1, chunk_target / (int64_t)line_bytes);
    ch = std::min(ch, batch);
    if (p->path == PATH_GENERIC || p->path == PATH_EXT) ch = std::min(ch, p->gbatch);
    // the generic paths share one work buffer between chunks: keep them on a single stream
    const bool generic = p->path == PATH_GENERIC || p->path == PATH_LARGE || p->path == PATH_EXT;
    // the O(n^2) kernel is not in-place safe: its chunks get separate input and output halves of the staging buffer
    const bool direct = p->path == PATH_DIRECT;
    const int64_t nchunks = (batch + ch - 1) / ch;
    const int nbuf = generic ? 1 : (int)std::min<int64_t>(HostPipe::NPIPE, nchunks);
    const size_t chunk_bytes = line_bytes * (size_t)ch;
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, chunk_bytes * (direct ? 2 : 1)));
    if (bounce_in || bounce_out) AFB_TRY(pipe->ensure_pinned(nbuf, chunk_bytes, bounce_in, bounce_out));
    struct Pending { char* dst = nullptr; int64_t nb = 0; bool busy = false; } pend[HostPipe::NPIPE];
    int32_t st = AFB_OK;
    // finish slot si: wait for its stream and, for a pageable destination, copy the finished chunk out of the ring
    auto drain = [&](int si) -> int32_t {
        if (!pend[si].busy) return AFB_OK;
        AFB_CUDA(cudaStreamSynchronize(pipe->stream[si]));
        if (bounce_out) copy_team().copy2d(pend[si].dst, pitch, (const char*)pipe->hout[si], line_bytes, line_bytes, (size_t)pend[si].nb);
        pend[si].busy = false;
        return AFB_OK;
    };
    int64_t c = 0;
    for (int64_t b0 = 0; b0 < batch && st == AFB_OK; b0 += ch, ++c) {
        const int si = (int)(c % nbuf);
        cudaStream_t s = pipe->stream[si];
        void* dbuf = pipe->buf[si];
        void* dout = direct ? (void*)((char*)dbuf + chunk_bytes) : dbuf;
        const int64_t nb = std::min(ch, batch - b0);
        const char* src = reinterpret_cast<const char*>(in) + pitch * (size_t)b0;
        char* dst = reinterpret_cast<char*>(out) + pitch * (size_t)b0;
        cudaError_t e;
        if (bounce_in || bounce_out) { st = drain(si); if (st != AFB_OK) break; }   // the slot's ring buffers are free again
        if (bounce_in) {
            copy_team().copy2d((char*)pipe->hin[si], line_bytes, src, pitch, line_bytes, (size_t)nb);
            e = cudaMemcpyAsync(dbuf, pipe->hin[si], line_bytes * (size_t)nb, cudaMemcpyHostToDevice, s);
        } else if (p->stride == n) {
            e = cudaMemcpyAsync(dbuf, src, line_bytes * (size_t)nb, cudaMemcpyHostToDevice, s);
        } else {
            e = cudaMemcpy2DAsync(dbuf, line_bytes, src, pitch, line_bytes, (size_t)nb, cudaMemcpyHostToDevice, s);
        }
        if (e != cudaSuccess) { st = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
        st = exec_1d_device(p, dbuf, dout, nb, n, n, s);
        if (st != AFB_OK) break;
        if (bounce_out) e = cudaMemcpyAsync(pipe->hout[si], dout, line_bytes * (size_t)nb, cudaMemcpyDeviceToHost, s);
        else if (p->stride == n) e = cudaMemcpyAsync(dst, dout, line_bytes * (size_t)nb, cudaMemcpyDeviceToHost, s);
        else e = cudaMemcpy2DAsync(dst, pitch, dout, line_bytes, line_bytes, (size_t)nb, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { st = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
        pend[si].dst = dst; pend[si].nb = nb; pend[si].busy = true;
    }
    // the pipe goes back to the pool only when its streams are idle (also on the error path); chunks drain in issue order
    for (int k = 0; k < nbuf; ++k) {
        const int si = (int)((c + k) % nbuf);
        if (st == AFB_OK) st = drain(si);
        else { cudaStreamSynchronize(pipe->stream[si]); cudaGetLastError(); }
    }
    return st;
}

static int32_t plan_create_1d(afb_plan_s** out, int32_t kind, int64_t n, int64_t batch, int64_t stride, uint32_t flags);

// the replica of a 1-D plan on device `dev` (current), able to run `batch` lines
static int32_t plan_replica(afb_plan_s* p, int dev, int64_t batch, afb_plan_s** out) {
    if (dev == p->dev) { *out = p; return AFB_OK; }
    for (size_t i = 0; i < p->replicas.size(); ++i)
        if (p->replica_devs[i] == dev && p->replicas[i]->batch >= batch) { *out = p->replicas[i]; return AFB_OK; }
    afb_plan_s* r = nullptr;
    AFB_TRY(plan_create_1d(&r, p->kind, p->n, batch, p->stride, p->flags));
    p->replicas.push_back(r);
    p->replica_devs.push_back(dev);
    *out = r;
    return AFB_OK;
}

extern "C" {

int32_t afb_plan_execute_host(afb_plan p, const void* in, void* out) {
    if (!p) return fail(AFB_BAD_ARG, "afb_plan_execute_host: null plan");
    if (p->path == PATH_EMPTY) return AFB_OK;
    if (!in || !out) return fail(AFB_BAD_ARG, "afb_plan_execute_host: null buffer");
    std::vector<int> devs;
    { std::lock_guard<std::mutex> lk(g_mu); devs = g_devlist; }
    if (p->path == PATH_TENSOR2D || p->path == PATH_PADUA) {
        if (p->path == PATH_TENSOR2D && devs.size() > 1 && mdev2d_supported(p->kind, p->n, p->n2, (int)devs.size(), p->flags)) {
            std::lock_guard<std::mutex> lk(p->mu);
            if (p->mdev2d && !mdev2d_matches(p->mdev2d, devs)) { mdev2d_destroy(p->mdev2d); p->mdev2d = nullptr; }   // afb_init changed the set
            if (!p->mdev2d) AFB_TRY(mdev2d_create(&p->mdev2d, p->kind, p->n, p->n2, devs));
            return mdev2d_execute_host(p->mdev2d, reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out));
        }
        DeviceGuard guard(p->dev);
        const size_t bytes = (size_t)p->elem * (size_t)(p->path == PATH_PADUA ? p->n : p->n * p->n2);
        void* d = nullptr;
        AFB_CUDA(cudaMalloc(&d, bytes));
        int32_t st = big_h2d(d, in, bytes);            // pageable caller arrays: through the pinned ring
        if (st == AFB_OK) st = afb_plan_execute_device(p, d, d, nullptr);
        if (st == AFB_OK) st = big_d2h(out, d, bytes);
        cudaFree(d);
        return st;
    }
    std::lock_guard<std::mutex> lk(p->mu);
    // several devices: contiguous shares of the batch, one host thread + one pipe per device (no collective)
    const int64_t total_bytes = (int64_t)p->elem * p->n * p->batch;
    if (devs.size() > 1 && p->batch >= (int64_t)devs.size() && total_bytes >= ((int64_t)8 << 20)) {
        const int64_t share = (p->batch + (int64_t)devs.size() - 1) / (int64_t)devs.size();
        return fanout(devs, p->batch, [&](int g, int64_t lo, int64_t hi) -> int32_t {
            afb_plan_s* r = nullptr;
            {
                static std::mutex rep_mu;
                std::lock_guard<std::mutex> rl(rep_mu);
                AFB_TRY(plan_replica(p, devs[(size_t)g], share, &r));
            }
            const size_t off = (size_t)p->elem * (size_t)(lo * p->stride);
            return host_pipeline_1d(r, reinterpret_cast<const char*>(in) + off, reinterpret_cast<char*>(out) + off, hi - lo);
        });
    }
    DeviceGuard guard(p->dev);
    return host_pipeline_1d(p, in, out, p->batch);
}

// ---- points -------------------------------------------------------------------------------------------
int32_t afb_chebpoints_device(double* d_out, int64_t n, double a, double b, void* stream) {
    if (n < 0) return fail(AFB_BAD_SIZE, "afb_chebpoints: n < 0");
    if (n == 0) return AFB_OK;
    if (!d_out) return fail(AFB_BAD_ARG, "afb_chebpoints: null buffer");
    AFB_TRY(ensure_device());
    return launch_chebpoints(d_out, n, a, b, stream);
}

}  // extern "C"

// generic "copy in, run, copy out" helper for the plain-function entry points.  Small device blocks are recycled through
// a process-wide cache: cudaMalloc / cudaFree per call cost more than the kernels of a scalar f(x) (69 us -> see DESIGN).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int dev = 0;
    ~DevBuf() { if (p) block_cache_put(p, cap, dev); }
    int32_t alloc(size_t bytes) { dev = current_device(); return block_cache_get(bytes ? bytes : 1, &p, &cap); }
    int32_t put(const void* h, size_t bytes) { return big_h2d(p, h, bytes); }
    int32_t get(void* h, size_t bytes) { return big_d2h(h, p, bytes); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// point sets / sample sets shard over afb_init's devices when there is enough work to pay for the host threads
static bool fan_points(int64_t P, int64_t work_per_point, std::vector<int>* devs) {
    { std::lock_guard<std::mutex> lk(g_mu); *devs = device_list(); }
    return devs->size() > 1 && P >= ((int64_t)1 << 20) && (double)P * (double)std::max<int64_t>(work_per_point, 1) >= 1e9;
}

// Large point sets through host pointers: chunks of 2^22 points ping-pong over the streams of a leased pipe -- H2D of chunk
// c+1 and D2H of chunk c-1 run under the kernel of chunk c, and the device never holds more than three chunks (a 10^9-point
// call would otherwise need 16 GB of device memory and serialise 16 GB of copies around 1.3 s of arithmetic).
// launch(d_in, d_out, np, stream) evaluates np points.
static constexpr int64_t POINT_CHUNK = (int64_t)1 << 22;
template <class Launch>
static int32_t host_points_pipeline(const double* in, double* out, int64_t P, Launch launch) {
    const int nbuf = (int)std::min<int64_t>(HostPipe::NPIPE, (P + POINT_CHUNK - 1) / POINT_CHUNK);
    const size_t chunk_bytes = sizeof(double) * (size_t)std::min(P, POINT_CHUNK);
    // pageable caller arrays go through the pinned ring (a pageable cudaMemcpyAsync is staged by the driver on one thread and
    // blocks the host, which serialises the chunks: for cheap kernels -- sampling -- the call was D2H bound)
    const bool big = sizeof(double) * (size_t)P >= ((size_t)4 << 20);
    const bool bounce_in = big && host_pageable(in), bounce_out = big && host_pageable(out);
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, 2 * chunk_bytes));   // input half, output half
    if (bounce_in || bounce_out) AFB_TRY(pipe->ensure_pinned(nbuf, chunk_bytes, bounce_in, bounce_out));
    struct Pending { double* dst = nullptr; int64_t np = 0; bool busy = false; } pend[HostPipe::NPIPE];
    auto drain = [&](int si) -> int32_t {
        if (!pend[si].busy) return AFB_OK;
        AFB_CUDA(cudaStreamSynchronize(pipe->stream[si]));
        const size_t nb = sizeof(double) * (size_t)pend[si].np;
        if (bounce_out) copy_team().copy2d((char*)pend[si].dst, nb, (const char*)pipe->hout[si], nb, nb, 1);
        pend[si].busy = false;
        return AFB_OK;
    };
    int32_t st = AFB_OK;
    int64_t c = 0;
    for (int64_t p0 = 0; p0 < P && st == AFB_OK; p0 += POINT_CHUNK, ++c) {
        const int si = (int)(c % nbuf);
        const int64_t np = std::min(POINT_CHUNK, P - p0);
        const size_t nb = sizeof(double) * (size_t)np;
        cudaStream_t s = pipe->stream[si];
        double* din = reinterpret_cast<double*>(pipe->buf[si]);
        double* dout = reinterpret_cast<double*>((char*)pipe->buf[si] + chunk_bytes);
        if (bounce_in || bounce_out) { st = drain(si); if (st != AFB_OK) break; }
        const void* hsrc = in + p0;
        if (bounce_in) { copy_team().copy2d((char*)pipe->hin[si], nb, (const char*)(in + p0), nb, nb, 1); hsrc = pipe->hin[si]; }
        cudaError_t e = cudaMemcpyAsync(din, hsrc, nb, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { st = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
        st = launch(din, dout, np, (void*)s);
        if (st != AFB_OK) break;
        e = cudaMemcpyAsync(bounce_out ? pipe->hout[si] : (void*)(out + p0), dout, nb, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { st = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
        pend[si].dst = out + p0; pend[si].np = np; pend[si].busy = true;
    }
    for (int k = 0; k < nbuf; ++k) {
        const int si = (int)((c + k) % nbuf);
        if (st == AFB_OK) st = drain(si);
        else { cudaStreamSynchronize(pipe->stream[si]); cudaGetLastError(); }
    }
    return st;
}

extern "C" {

int32_t afb_chebpoints(double* out, int64_t n, double a, double b) {
    if (n < 0) return fail(AFB_BAD_SIZE, "afb_chebpoints: n < 0");
    if (n == 0) return AFB_OK;
    if (!out) return fail(AFB_BAD_ARG, "afb_chebpoints: null buffer");
    AFB_TRY(ensure_device());
    DevBuf d;
    AFB_TRY(d.alloc(sizeof(double) * (size_t)n));
    AFB_TRY(launch_chebpoints(d.as<double>(), n, a, b, nullptr));
    return d.get(out, sizeof(double) * (size_t)n);
}

int32_t afb_fourierpoints(double* out, int64_t n, double a, double b) {
    if (n < 0) return fail(AFB_BAD_SIZE, "afb_fourierpoints: n < 0");
    if (n && !out) return fail(AFB_BAD_ARG, "afb_fourierpoints: null buffer");
    for (int64_t j = 0; j < n; ++j) out[j] = a + (b - a) * (double)j / (double)n;  // O(n) host arithmetic, no GPU analogue
    return AFB_OK;
}

int32_t afb_paduapoints(double* xy, int64_t N, double ax, double bx, double ay, double by, int64_t* np_out) {
    if (N < 0) return fail(AFB_BAD_SIZE, "afb_paduapoints: N < 0");
    const int64_t d = padua_degree(N), Np = (N == 0) ? 0 : (d + 1) * (d + 2) / 2;
    if (np_out) *np_out = Np;
    if (!xy || Np == 0) return AFB_OK;
    if (d == 0) {  // upstream evaluates -cospi(0/0) here; the single point is the lower-left corner
        xy[0] = ax; xy[1] = ay;
        return AFB_OK;
    }
    if ((d + 2) * (d + 1) >= ((int64_t)1 << 31)) return fail(AFB_BAD_SIZE, "afb_paduapoints: degree too large");
    AFB_TRY(ensure_device());
    int32_t* idx = nullptr;
    AFB_TRY(upload_idx(padua_grid_index(d), &idx));
    DevBuf dev;
    int32_t st = dev.alloc(sizeof(double) * (size_t)(2 * Np));
    if (st == AFB_OK) st = launch_paduapoints(dev.as<double>(), idx, Np, d, ax, bx, ay, by, nullptr);
    if (st == AFB_OK) st = dev.get(xy, sizeof(double) * (size_t)(2 * Np));
    cudaFree(idx);
    return st;
}

// ---- Clenshaw ----------------------------------------------------------------------------------------
int32_t afb_clenshaw_cheb_device(const double* d_c, int64_t N, const double* d_x, double* d_y, int64_t P, void* stream) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !d_c) || !d_x || !d_y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb: null buffer");
    AFB_TRY(ensure_device());
    return launch_clenshaw_vec(d_c, N, d_x, d_y, P, stream);
}
int32_t afb_clenshaw_cheb(const double* c, int64_t N, const double* x, double* y, int64_t P) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !c) || !x || !y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb: null buffer");
    AFB_TRY(ensure_device());
    std::vector<int> devs;
    if (fan_points(P, N, &devs))
        return fanout(devs, P, [&](int, int64_t lo, int64_t hi) { return afb_clenshaw_cheb(c, N, x + lo, y + lo, hi - lo); });
    DevBuf dc, dx, dy;
    AFB_TRY(dc.alloc(8 * (size_t)N));
    AFB_TRY(dc.put(c, 8 * (size_t)N));
    if (P > POINT_CHUNK)
        return host_points_pipeline(x, y, P, [&](const double* din, double* dout, int64_t np, void* s) {
            return launch_clenshaw_vec(dc.as<double>(), N, din, dout, np, s);
        });
    AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)P));
    AFB_TRY(dx.put(x, 8 * (size_t)P));
    AFB_TRY(launch_clenshaw_vec(dc.as<double>(), N, dx.as<double>(), dy.as<double>(), P, nullptr));
    return dy.get(y, 8 * (size_t)P);
}

int32_t afb_clenshaw_rec_device(const double* d_c, int64_t N, const double* d_A, const double* d_B, const double* d_C,
                                const double* d_x, double* d_y, int64_t P, void* stream) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_rec: negative size");
    if (P == 0) return AFB_OK;
    if ((N && (!d_c || !d_A || !d_B || !d_C)) || !d_x || !d_y) return fail(AFB_BAD_ARG, "afb_clenshaw_rec: null buffer");
    AFB_TRY(ensure_device());
    return launch_clenshaw_rec(d_c, N, d_A, d_B, d_C, d_x, d_y, P, stream);
}
int32_t afb_clenshaw_rec(const double* c, int64_t N, const double* A, const double* B, const double* C, const double* x,
                         double* y, int64_t P) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_rec: negative size");
    if (P == 0) return AFB_OK;
    if ((N && (!c || !A || !B || !C)) || !x || !y) return fail(AFB_BAD_ARG, "afb_clenshaw_rec: null buffer");
    AFB_TRY(ensure_device());
    std::vector<int> devs;
    if (fan_points(P, N, &devs))
        return fanout(devs, P, [&](int, int64_t lo, int64_t hi) { return afb_clenshaw_rec(c, N, A, B, C, x + lo, y + lo, hi - lo); });
    DevBuf dc, dA, dB, dC, dx, dy;
    AFB_TRY(dc.alloc(8 * (size_t)N)); AFB_TRY(dA.alloc(8 * (size_t)N)); AFB_TRY(dB.alloc(8 * (size_t)N));
    AFB_TRY(dC.alloc(8 * (size_t)(N + 1)));
    AFB_TRY(dc.put(c, 8 * (size_t)N)); AFB_TRY(dA.put(A, 8 * (size_t)N)); AFB_TRY(dB.put(B, 8 * (size_t)N));
    if (N) AFB_TRY(dC.put(C, 8 * (size_t)(N + 1)));
    if (P > POINT_CHUNK)
        return host_points_pipeline(x, y, P, [&](const double* din, double* dout, int64_t np, void* s) {
            return launch_clenshaw_rec(dc.as<double>(), N, dA.as<double>(), dB.as<double>(), dC.as<double>(), din, dout, np, s);
        });
    AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)P));
    AFB_TRY(dx.put(x, 8 * (size_t)P));
    AFB_TRY(launch_clenshaw_rec(dc.as<double>(), N, dA.as<double>(), dB.as<double>(), dC.as<double>(), dx.as<double>(),
                                dy.as<double>(), P, nullptr));
    return dy.get(y, 8 * (size_t)P);
}

int32_t afb_clenshaw_cheb_cols_device(const double* d_C, int64_t N, int64_t ld, const double* d_x, double* d_y,
                                      int64_t P, void* stream) {
    if (N < 0 || P < 0 || ld < N) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb_cols: bad sizes (need ld >= N >= 0)");
    if (P == 0) return AFB_OK;
    if ((N && !d_C) || !d_x || !d_y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb_cols: null buffer");
    AFB_TRY(ensure_device());
    return launch_cols(false, d_C, N, ld, d_x, d_y, P, 0, stream);
}
int32_t afb_clenshaw_cheb_cols(const double* C, int64_t N, int64_t ld, const double* x, double* y, int64_t P) {
    if (N < 0 || P < 0 || ld < N) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb_cols: bad sizes (need ld >= N >= 0)");
    if (P == 0) return AFB_OK;
    if ((N && !C) || !x || !y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb_cols: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dc, dx, dy;
    const size_t cb = 8 * (size_t)(ld * P);
    AFB_TRY(dc.alloc(cb)); AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)P));
    AFB_TRY(dc.put(C, cb)); AFB_TRY(dx.put(x, 8 * (size_t)P));
    AFB_TRY(launch_cols(false, dc.as<double>(), N, ld, dx.as<double>(), dy.as<double>(), P, 0, nullptr));
    return dy.get(y, 8 * (size_t)P);
}

int32_t afb_clenshaw_cheb_multi_device(const double* d_C, int64_t N, int64_t M, const double* d_x, double* d_Y,
                                       int64_t P, void* stream) {
    if (N < 0 || M < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb_multi: negative size");
    if (P == 0 || M == 0) return AFB_OK;
    if ((N && !d_C) || !d_x || !d_Y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb_multi: null buffer");
    AFB_TRY(ensure_device());
    return launch_multi(d_C, N, M, d_x, d_Y, P, stream);
}
int32_t afb_clenshaw_cheb_multi(const double* C, int64_t N, int64_t M, const double* x, double* Y, int64_t P) {
    if (N < 0 || M < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb_multi: negative size");
    if (P == 0 || M == 0) return AFB_OK;
    if ((N && !C) || !x || !Y) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb_multi: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dc, dx, dy;
    AFB_TRY(dc.alloc(8 * (size_t)(N * M))); AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)(M * P)));
    AFB_TRY(dc.put(C, 8 * (size_t)(N * M))); AFB_TRY(dx.put(x, 8 * (size_t)P));
    AFB_TRY(launch_multi(dc.as<double>(), N, M, dx.as<double>(), dy.as<double>(), P, nullptr));
    return dy.get(Y, 8 * (size_t)(M * P));
}

// ---- sampling ------------------------------------------------------------------------------------------
int32_t afb_cheb_normcumsum_device(double* d_C, int64_t N, int64_t ncols, int64_t ld, int32_t grow, void* stream) {
    if (N < 0 || ncols < 0 || ld < N + (grow ? 1 : 0))
        return fail(AFB_BAD_SIZE, "afb_cheb_normcumsum: bad sizes (need ld >= N + grow)");
    if (ncols == 0) return AFB_OK;
    if (!d_C) return fail(AFB_BAD_ARG, "afb_cheb_normcumsum: null buffer");
    AFB_TRY(ensure_device());
    return launch_normcumsum(d_C, N, ncols, ld, grow, stream);
}
int32_t afb_cheb_normcumsum(double* C, int64_t N, int64_t ncols, int64_t ld, int32_t grow) {
    if (N < 0 || ncols < 0 || ld < N + (grow ? 1 : 0))
        return fail(AFB_BAD_SIZE, "afb_cheb_normcumsum: bad sizes (need ld >= N + grow)");
    if (ncols == 0) return AFB_OK;
    if (!C) return fail(AFB_BAD_ARG, "afb_cheb_normcumsum: null buffer");
    AFB_TRY(ensure_device());
    DevBuf d;
    const size_t bytes = 8 * (size_t)(ld * ncols);
    AFB_TRY(d.alloc(bytes)); AFB_TRY(d.put(C, bytes));
    AFB_TRY(launch_normcumsum(d.as<double>(), N, ncols, ld, grow, nullptr));
    return d.get(C, bytes);
}

int32_t afb_cheb_bisect_device(const double* d_c, int64_t N, const double* d_u, double* d_out, int64_t P,
                               int32_t numits, double a, double b, void* stream) {
    if (N < 0 || P < 0 || numits < 0) return fail(AFB_BAD_SIZE, "afb_cheb_bisect: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !d_c) || !d_u || !d_out) return fail(AFB_BAD_ARG, "afb_cheb_bisect: null buffer");
    AFB_TRY(ensure_device());
    return launch_bisect_vec(d_c, N, d_u, false, 0, 0, d_out, P, numits, a, b, stream);
}
static int32_t bisect_host(const double* c, int64_t N, const double* u, double* out, int64_t P, int32_t numits,
                           double a, double b) {
    if (N < 0 || P < 0 || numits < 0) return fail(AFB_BAD_SIZE, "afb_cheb_bisect: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !c) || !u || !out) return fail(AFB_BAD_ARG, "afb_cheb_bisect: null buffer");
    AFB_TRY(ensure_device());
    std::vector<int> devs;
    if (fan_points(P, (int64_t)numits * N, &devs))
        return fanout(devs, P, [&](int, int64_t lo, int64_t hi) { return bisect_host(c, N, u + lo, out + lo, hi - lo, numits, a, b); });
    DevBuf dc, du, dout;
    AFB_TRY(dc.alloc(8 * (size_t)N));
    AFB_TRY(dc.put(c, 8 * (size_t)N));
    if (P > POINT_CHUNK)
        return host_points_pipeline(u, out, P, [&](const double* din, double* dres, int64_t np, void* s) {
            return launch_bisect_vec(dc.as<double>(), N, din, false, 0, 0, dres, np, numits, a, b, s);
        });
    AFB_TRY(du.alloc(8 * (size_t)P)); AFB_TRY(dout.alloc(8 * (size_t)P));
    AFB_TRY(du.put(u, 8 * (size_t)P));
    AFB_TRY(launch_bisect_vec(dc.as<double>(), N, du.as<double>(), false, 0, 0, dout.as<double>(), P, numits, a, b, nullptr));
    return dout.get(out, 8 * (size_t)P);
}
int32_t afb_cheb_bisect(const double* c, int64_t N, const double* u, double* out, int64_t P, int32_t numits) {
    return bisect_host(c, N, u, out, P, numits, -1.0, 1.0);
}
int32_t afb_samplecdf_cheb(const double* cdf_c, int64_t N, double a, double b, const double* u, double* out, int64_t P) {
    return bisect_host(cdf_c, N, u, out, P, 47, a, b);
}

int32_t afb_cheb_bisect_cols_device(const double* d_C, int64_t N, int64_t ld, const double* d_u, double* d_out,
                                    int64_t P, int32_t numits, void* stream) {
    if (N < 0 || P < 0 || ld < N || numits < 0) return fail(AFB_BAD_SIZE, "afb_cheb_bisect_cols: bad sizes");
    if (P == 0) return AFB_OK;
    if ((N && !d_C) || !d_u || !d_out) return fail(AFB_BAD_ARG, "afb_cheb_bisect_cols: null buffer");
    AFB_TRY(ensure_device());
    return launch_cols(true, d_C, N, ld, d_u, d_out, P, numits, stream);
}
int32_t afb_cheb_bisect_cols(const double* C, int64_t N, int64_t ld, const double* u, double* out, int64_t P,
                             int32_t numits) {
    if (N < 0 || P < 0 || ld < N || numits < 0) return fail(AFB_BAD_SIZE, "afb_cheb_bisect_cols: bad sizes");
    if (P == 0) return AFB_OK;
    if ((N && !C) || !u || !out) return fail(AFB_BAD_ARG, "afb_cheb_bisect_cols: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dc, du, dout;
    const size_t cb = 8 * (size_t)(ld * P);
    AFB_TRY(dc.alloc(cb)); AFB_TRY(du.alloc(8 * (size_t)P)); AFB_TRY(dout.alloc(8 * (size_t)P));
    AFB_TRY(dc.put(C, cb)); AFB_TRY(du.put(u, 8 * (size_t)P));
    AFB_TRY(launch_cols(true, dc.as<double>(), N, ld, du.as<double>(), dout.as<double>(), P, numits, nullptr));
    return dout.get(out, 8 * (size_t)P);
}

int32_t afb_sample_cheb_device(const double* d_cdf_c, int64_t N, double a, double b, uint64_t seed, uint64_t offset,
                               double* d_out, int64_t P, void* stream) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_sample_cheb: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !d_cdf_c) || !d_out) return fail(AFB_BAD_ARG, "afb_sample_cheb: null buffer");
    AFB_TRY(ensure_device());
    return launch_bisect_vec(d_cdf_c, N, nullptr, true, seed, offset, d_out, P, 47, a, b, stream);
}
// host-pointer sampler with the uniforms drawn on the device (Philox4x32-10, sample i <- counter offset + i): chunks of
// 2^23 samples ping-pong between two streams so the kernel of chunk i+1 runs under the D2H copy of chunk i
static int32_t sample_host_one(const double* cdf_c, int64_t N, double a, double b, uint64_t seed, uint64_t offset, double* out,
                               int64_t P) {
    DevBuf dc;
    AFB_TRY(dc.alloc(8 * (size_t)N));
    AFB_TRY(dc.put(cdf_c, 8 * (size_t)N));
    // a pageable destination (a Julia / NumPy array) is drained through the pinned ring: the driver's own pageable D2H runs at
    // a fraction of PCIe speed on one thread and blocks the host, which made sample(f, 10^9) copy bound
    const bool bounce = sizeof(double) * (size_t)P >= ((size_t)4 << 20) && host_pageable(out);
    const int64_t chunk = std::min<int64_t>(P, bounce ? POINT_CHUNK : ((int64_t)1 << 23));
    const int nbuf = (int)std::min<int64_t>(bounce ? HostPipe::NPIPE : 2, (P + chunk - 1) / chunk);
    PipeLease pipe;
    AFB_TRY(pipe->ensure(nbuf, sizeof(double) * (size_t)chunk));
    if (bounce) AFB_TRY(pipe->ensure_pinned(nbuf, sizeof(double) * (size_t)chunk, false, true));
    struct Pending { double* dst = nullptr; int64_t np = 0; bool busy = false; } pend[HostPipe::NPIPE];
    auto drain = [&](int si) -> int32_t {
        if (!pend[si].busy) return AFB_OK;
        AFB_CUDA(cudaStreamSynchronize(pipe->stream[si]));
        const size_t nb = sizeof(double) * (size_t)pend[si].np;
        if (bounce) copy_team().copy2d((char*)pend[si].dst, nb, (const char*)pipe->hout[si], nb, nb, 1);
        pend[si].busy = false;
        return AFB_OK;
    };
    int64_t c = 0;
    int32_t st = AFB_OK;
    for (int64_t p0 = 0; p0 < P && st == AFB_OK; p0 += chunk, ++c) {
        const int si = (int)(c % nbuf);
        const int64_t np = std::min(chunk, P - p0);
        st = drain(si);                                // the slot's device and ring buffers are free again
        if (st != AFB_OK) break;
        double* dbuf = reinterpret_cast<double*>(pipe->buf[si]);
        st = launch_bisect_vec(dc.as<double>(), N, nullptr, true, seed, offset + (uint64_t)p0, dbuf, np, 47, a, b, pipe->stream[si]);
        if (st != AFB_OK) break;
        cudaError_t e = cudaMemcpyAsync(bounce ? pipe->hout[si] : (void*)(out + p0), dbuf, sizeof(double) * (size_t)np,
                                        cudaMemcpyDeviceToHost, pipe->stream[si]);
        if (e != cudaSuccess) { st = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
        pend[si].dst = out + p0; pend[si].np = np; pend[si].busy = true;
    }
    for (int k = 0; k < nbuf; ++k) {
        const int si = (int)((c + k) % nbuf);
        if (st == AFB_OK) st = drain(si);
        else { cudaStreamSynchronize(pipe->stream[si]); cudaGetLastError(); }
    }
    return st;
}
int32_t afb_sample_cheb(const double* cdf_c, int64_t N, double a, double b, uint64_t seed, double* out, int64_t P) {
    if (N < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_sample_cheb: negative size");
    if (P == 0) return AFB_OK;
    if ((N && !cdf_c) || !out) return fail(AFB_BAD_ARG, "afb_sample_cheb: null buffer");
    AFB_TRY(ensure_device());
    std::vector<int> devs;
    if (fan_points(P, 47 * N, &devs))   // sample i always uses Philox counter i: the stream does not depend on the device count
        return fanout(devs, P, [&](int, int64_t lo, int64_t hi) { return sample_host_one(cdf_c, N, a, b, seed, (uint64_t)lo, out + lo, hi - lo); });
    return sample_host_one(cdf_c, N, a, b, seed, 0, out, P);
}
int32_t afb_philox_uniform_device(uint64_t seed, uint64_t offset, double* d_out, int64_t P, void* stream) {
    if (P < 0) return fail(AFB_BAD_SIZE, "afb_philox_uniform: negative size");
    if (P == 0) return AFB_OK;
    if (!d_out) return fail(AFB_BAD_ARG, "afb_philox_uniform: null buffer");
    AFB_TRY(ensure_device());
    return launch_philox(seed, offset, d_out, P, stream);
}


// ---- piecewise Chebyshev Funs: evaluation and the generic bisection of sample.jl:8-36 -----------------------------
static int32_t piecewise_host(const double* coef, const int64_t* offsets, const double* breaks, int64_t np, const double* xu,
                              double* out, int64_t P, bool bisect, int32_t numits) {
    if (np < 1 || P < 0 || numits < 0) return fail(AFB_BAD_SIZE, "afb_piecewise: need npieces >= 1, P >= 0, numits >= 0");
    if (P == 0) return AFB_OK;
    if (!offsets || !breaks || !xu || !out) return fail(AFB_BAD_ARG, "afb_piecewise: null buffer");
    if (offsets[0] != 0) return fail(AFB_BAD_SIZE, "afb_piecewise: offsets[0] must be 0");
    for (int64_t k = 0; k < np; ++k) {
        if (offsets[k + 1] < offsets[k]) return fail(AFB_BAD_SIZE, "afb_piecewise: offsets must be non-decreasing");
        if (!(breaks[k] < breaks[k + 1])) return fail(AFB_BAD_SIZE, "afb_piecewise: break points must be increasing");
    }
    const int64_t total = offsets[np];
    if (total >= ((int64_t)1 << 31)) return fail(AFB_BAD_SIZE, "afb_piecewise: too many coefficients");
    if (total && !coef) return fail(AFB_BAD_ARG, "afb_piecewise: null coefficients");
    AFB_TRY(ensure_device());
    std::vector<int32_t> off32((size_t)np + 1);
    for (int64_t k = 0; k <= np; ++k) off32[(size_t)k] = (int32_t)offsets[k];
    DevBuf dc, doff, dbrk, dxu, dout;
    AFB_TRY(dc.alloc(8 * (size_t)total)); AFB_TRY(doff.alloc(4 * (size_t)(np + 1))); AFB_TRY(dbrk.alloc(8 * (size_t)(np + 1)));
    AFB_TRY(dxu.alloc(8 * (size_t)P)); AFB_TRY(dout.alloc(8 * (size_t)P));
    AFB_TRY(dc.put(coef, 8 * (size_t)total)); AFB_TRY(doff.put(off32.data(), 4 * (size_t)(np + 1)));
    AFB_TRY(dbrk.put(breaks, 8 * (size_t)(np + 1))); AFB_TRY(dxu.put(xu, 8 * (size_t)P));
    AFB_TRY(launch_piecewise(dc.as<double>(), doff.as<int32_t>(), dbrk.as<double>(), np, total, dxu.as<double>(), dout.as<double>(),
                             P, bisect, numits, nullptr));
    return dout.get(out, 8 * (size_t)P);
}
int32_t afb_piecewise_clenshaw(const double* coef, const int64_t* offsets, const double* breaks, int64_t npieces, const double* x,
                               double* y, int64_t P) {
    return piecewise_host(coef, offsets, breaks, npieces, x, y, P, false, 0);
}
int32_t afb_piecewise_bisect(const double* coef, const int64_t* offsets, const double* breaks, int64_t npieces, const double* u,
                             double* out, int64_t P, int32_t numits) {
    return piecewise_host(coef, offsets, breaks, npieces, u, out, P, true, numits);
}

// ---- bivariate evaluation of a Chebyshev x Chebyshev coefficient matrix --------------------------------------
int32_t afb_clenshaw_cheb2d_device(const double* d_C, int64_t n1, int64_t n2, const double* d_x, const double* d_y,
                                   double* d_out, int64_t P, void* stream) {
    if (n1 < 0 || n2 < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb2d: negative size");
    if (P == 0) return AFB_OK;
    if ((n1 * n2 && !d_C) || !d_x || !d_y || !d_out) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb2d: null buffer");
    AFB_TRY(ensure_device());
    return launch_clenshaw2d(d_C, n1, n2, d_x, d_y, d_out, P, stream);
}
int32_t afb_clenshaw_cheb2d(const double* C, int64_t n1, int64_t n2, const double* x, const double* y, double* out,
                            int64_t P) {
    if (n1 < 0 || n2 < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_cheb2d: negative size");
    if (P == 0) return AFB_OK;
    if ((n1 * n2 && !C) || !x || !y || !out) return fail(AFB_BAD_ARG, "afb_clenshaw_cheb2d: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dc, dx, dy, dout;
    AFB_TRY(dc.alloc(8 * (size_t)(n1 * n2))); AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)P));
    AFB_TRY(dout.alloc(8 * (size_t)P));
    AFB_TRY(dc.put(C, 8 * (size_t)(n1 * n2))); AFB_TRY(dx.put(x, 8 * (size_t)P)); AFB_TRY(dy.put(y, 8 * (size_t)P));
    AFB_TRY(launch_clenshaw2d(dc.as<double>(), n1, n2, dx.as<double>(), dy.as<double>(), dout.as<double>(), P, nullptr));
    return dout.get(out, 8 * (size_t)P);
}

int32_t afb_clenshaw_fourier_cheb2d_device(const double* d_C, int64_t n1, int64_t n2, const double* d_t, const double* d_y,
                                           double* d_out, int64_t P, void* stream) {
    if (n1 < 0 || n2 < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_fourier_cheb2d: negative size");
    if (P == 0) return AFB_OK;
    if ((n1 * n2 && !d_C) || !d_t || !d_y || !d_out) return fail(AFB_BAD_ARG, "afb_clenshaw_fourier_cheb2d: null buffer");
    AFB_TRY(ensure_device());
    return launch_fourier_cheb2d(d_C, n1, n2, d_t, d_y, d_out, P, stream);
}
int32_t afb_clenshaw_fourier_cheb2d(const double* C, int64_t n1, int64_t n2, const double* t, const double* y, double* out,
                                    int64_t P) {
    if (n1 < 0 || n2 < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_clenshaw_fourier_cheb2d: negative size");
    if (P == 0) return AFB_OK;
    if ((n1 * n2 && !C) || !t || !y || !out) return fail(AFB_BAD_ARG, "afb_clenshaw_fourier_cheb2d: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dc, dx, dy, dout;
    AFB_TRY(dc.alloc(8 * (size_t)(n1 * n2))); AFB_TRY(dx.alloc(8 * (size_t)P)); AFB_TRY(dy.alloc(8 * (size_t)P));
    AFB_TRY(dout.alloc(8 * (size_t)P));
    AFB_TRY(dc.put(C, 8 * (size_t)(n1 * n2))); AFB_TRY(dx.put(t, 8 * (size_t)P)); AFB_TRY(dy.put(y, 8 * (size_t)P));
    AFB_TRY(launch_fourier_cheb2d(dc.as<double>(), n1, n2, dx.as<double>(), dy.as<double>(), dout.as<double>(), P, nullptr));
    return dout.get(out, 8 * (size_t)P);
}

// ---- 2-D sampling pieces (src/Extras/sample.jl:205-214) ---------------------------------------------------
int32_t afb_sample2d_cheb_lowrank_device(const double* d_A, int64_t NA, const double* d_CB, int64_t N, int64_t R,
                                         double ya, double yb, double xa, double xb, const double* d_ry,
                                         const double* d_u, double* d_rx, int64_t P, void* stream) {
    if (NA < 0 || N < 0 || R < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_sample2d_cheb_lowrank: negative size");
    if (P == 0) return AFB_OK;
    if ((NA * R && !d_A) || (N * R && !d_CB) || !d_ry || !d_u || !d_rx)
        return fail(AFB_BAD_ARG, "afb_sample2d_cheb_lowrank: null buffer");
    AFB_TRY(ensure_device());
    return launch_sample2d(d_A, NA, d_CB, N, R, ya, yb, xa, xb, d_ry, d_u, d_rx, P, 47, stream);
}
int32_t afb_sample2d_cheb_lowrank(const double* A, int64_t NA, const double* CB, int64_t N, int64_t R, double ya,
                                  double yb, double xa, double xb, const double* ry, const double* u, double* rx,
                                  int64_t P) {
    if (NA < 0 || N < 0 || R < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_sample2d_cheb_lowrank: negative size");
    if (P == 0) return AFB_OK;
    if ((NA * R && !A) || (N * R && !CB) || !ry || !u || !rx) return fail(AFB_BAD_ARG, "afb_sample2d_cheb_lowrank: null buffer");
    AFB_TRY(ensure_device());
    DevBuf dA, dCB, dry, du, drx;
    AFB_TRY(dA.alloc(8 * (size_t)(NA * R))); AFB_TRY(dCB.alloc(8 * (size_t)(N * R)));
    AFB_TRY(dry.alloc(8 * (size_t)P)); AFB_TRY(du.alloc(8 * (size_t)P)); AFB_TRY(drx.alloc(8 * (size_t)P));
    AFB_TRY(dA.put(A, 8 * (size_t)(NA * R))); AFB_TRY(dCB.put(CB, 8 * (size_t)(N * R)));
    AFB_TRY(dry.put(ry, 8 * (size_t)P)); AFB_TRY(du.put(u, 8 * (size_t)P));
    AFB_TRY(launch_sample2d(dA.as<double>(), NA, dCB.as<double>(), N, R, ya, yb, xa, xb, dry.as<double>(), du.as<double>(),
                            drx.as<double>(), P, 47, nullptr));
    return drx.get(rx, 8 * (size_t)P);
}

// AB = CB * fA  (N x R times R x P, column-major; the dgemm of src/Extras/sample.jl:210)
int32_t afb_dgemm_nn_device(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P,
                            void* stream) {
    if (N < 0 || R < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_dgemm_nn: negative size");
    if (N * P == 0) return AFB_OK;
    if ((N * R && !d_CB) || (R * P && !d_fA) || !d_AB) return fail(AFB_BAD_ARG, "afb_dgemm_nn: null buffer");
    AFB_TRY(ensure_device());
    return launch_gemm_nn(d_CB, d_fA, d_AB, N, R, P, stream);   // bit-exact FMA order; afb_dgemm_nn_dmma_device: tensor cores
}
int32_t afb_dgemm_nn(const double* CB, const double* fA, double* AB, int64_t N, int64_t R, int64_t P) {
    if (N < 0 || R < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_dgemm_nn: negative size");
    if (N * P == 0) return AFB_OK;
    if ((N * R && !CB) || (R * P && !fA) || !AB) return fail(AFB_BAD_ARG, "afb_dgemm_nn: null buffer");
    AFB_TRY(ensure_device());
    DevBuf a, b, c;
    AFB_TRY(a.alloc(8 * (size_t)(N * R))); AFB_TRY(b.alloc(8 * (size_t)(R * P))); AFB_TRY(c.alloc(8 * (size_t)(N * P)));
    AFB_TRY(a.put(CB, 8 * (size_t)(N * R))); AFB_TRY(b.put(fA, 8 * (size_t)(R * P)));
    AFB_TRY(launch_gemm_nn(a.as<double>(), b.as<double>(), c.as<double>(), N, R, P, nullptr));
    return c.get(AB, 8 * (size_t)(N * P));
}


// measurement variant of afb_dgemm_nn_device on the FP64 tensor cores (DMMA); agrees with it to rounding, not bit for bit
int32_t afb_dgemm_nn_dmma_device(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P,
                                 void* stream) {
    if (N < 0 || R < 0 || P < 0) return fail(AFB_BAD_SIZE, "afb_dgemm_nn_dmma: negative size");
    if (N * P == 0) return AFB_OK;
    if ((N * R && !d_CB) || (R * P && !d_fA) || !d_AB) return fail(AFB_BAD_ARG, "afb_dgemm_nn_dmma: null buffer");
    AFB_TRY(ensure_device());
    return launch_gemm_nn_dmma(d_CB, d_fA, d_AB, N, R, P, stream);
}

// measurement helper for bench.py (not a reference entry point): see fp64_probe_kernel
int32_t afb_fp64_probe_device(int32_t mode, int32_t iters, int32_t blocks, double* d_out, void* stream) {
    if (mode < 0 || mode > 4 || iters < 0 || blocks <= 0 || !d_out) return fail(AFB_BAD_ARG, "afb_fp64_probe: bad argument");
    AFB_TRY(ensure_device());
    return launch_fp64_probe(mode, iters, blocks, d_out, stream);
}

}  // extern "C"
