// Kernel-source emulator shim (DEVELOPER TEST TOOL, NOT A PRODUCT PATH).
//
// When the .cu sources are compiled by g++ with -DAFB_HOST_EMU (tests/emu/Makefile), this header
// stands in for the CUDA runtime so that the *same kernel source* can be single-stepped on a CPU:
// every CUDA thread of a block is a ucontext fiber, __syncthreads() yields to a round-robin
// scheduler, warp shuffles go through a per-block exchange buffer.  It exists because the build
// container has no GPU: index arithmetic of the FFT passes is unit-tested here first
// (tests/test_emulated_kernels.py) and then validated on a B200.  The shipped library
// libapproxfun_b200.so is built by nvcc only, never contains this code, and the Python/Julia host
// layers refuse to load anything else (no CPU fallback).
#pragma once
#ifdef AFB_HOST_EMU

#include <ucontext.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2,
       cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; int major; int minor; size_t totalGlobalMem; char name[64]; };

namespace afb_emu {
struct Block {
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stacks;
    std::vector<char> done;
    ucontext_t sched;
    int cur = 0;
    std::vector<unsigned char> smem;
    std::vector<uint64_t> xchg;
    std::function<void()> body;
};
extern thread_local Block* g_blk;
extern thread_local uint3 g_tid, g_bid;
extern thread_local dim3 g_bdim, g_gdim;
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void yield_barrier();
}  // namespace afb_emu

#define threadIdx (afb_emu::g_tid)
#define blockIdx (afb_emu::g_bid)
#define blockDim (afb_emu::g_bdim)
#define gridDim (afb_emu::g_gdim)

static inline void __syncthreads() { afb_emu::yield_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}

template <class T>
static inline T afb_emu_shfl(T v, int src_lane_of_warp_fn(int lane, int arg), int arg) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    afb_emu::Block* b = afb_emu::g_blk;
    int tid = threadIdx.x;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    b->xchg[tid] = raw;
    afb_emu::yield_barrier();
    int lane = tid & 31, base = tid & ~31;
    int src = src_lane_of_warp_fn(lane, arg);
    if (src < 0 || src > 31 || base + src >= (int)blockDim.x) src = lane;
    uint64_t got = b->xchg[base + src];
    afb_emu::yield_barrier();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
static inline int afb_emu_xor(int lane, int m) { return lane ^ m; }
static inline int afb_emu_idx(int, int s) { return s; }
static inline int afb_emu_down(int lane, int d) { return lane + d; }
static inline int afb_emu_up(int lane, int d) { return lane - d; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return afb_emu_shfl(v, afb_emu_xor, m); }
template <class T> static inline T __shfl_sync(unsigned, T v, int s) { return afb_emu_shfl(v, afb_emu_idx, s); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return afb_emu_shfl(v, afb_emu_down, d); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return afb_emu_shfl(v, afb_emu_up, d); }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double sinpi(double x) {
    double k = std::nearbyint(x), r = x - k;
    double s = std::sin(M_PI * r);
    if (r == 0.0) s = 0.0;
    if (std::fabs(r) == 0.5) s = r > 0 ? 1.0 : -1.0;
    return (std::fmod(k, 2.0) == 0.0) ? s : -s;
}

using std::isinf;
static inline double cospi(double x) { return sinpi(0.5 - std::fabs(x)); }
static inline void sincospi(double x, double* s, double* c) { *s = sinpi(x); *c = sinpi(x + 0.5); }

// --- runtime API stand-ins (host memory is "device" memory) -------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
enum { cudaHostRegisterPortable = 1 };
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k) { return cudaMemcpyAsync(d, s, n, k); }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind,
                                            cudaStream_t = 0) {
    for (size_t r = 0; r < h; ++r) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return 0;
}
static inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) {
    return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k);
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; p->major = 10; p->minor = 0;
    p->totalGlobalMem = (size_t)8 << 30; std::strcpy(p->name, "AFB host emulator"); return 0;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define AFB_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                      \
        (void)(stream);                                                                       \
        afb_emu::run_grid(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); }); \
    } while (0)
#define AFB_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(afb_emu::g_blk->smem.data())

#endif  // AFB_HOST_EMU
