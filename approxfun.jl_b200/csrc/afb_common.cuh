// Shared declarations for libapproxfun_b200.so (sm_100a).  See include/approxfun_b200.h for the ABI.
#pragma once

#ifdef AFB_HOST_EMU
#include "afb_emu.h"
#define AFB_HD
#else
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#define AFB_HD __host__ __device__
#define AFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define AFB_DYN_SMEM(T, name)                                          \
    extern __shared__ __align__(16) unsigned char afb_dyn_smem_raw[];  \
    T* name = reinterpret_cast<T*>(afb_dyn_smem_raw)
#endif

#include <cstddef>
#include <cstring>
#include <cstdint>
#include <string>

#include "../../include/approxfun_b200.h"

namespace afb {

// ---- status / error plumbing (never throws across the ABI) -------------------------------------
void set_error(int32_t code, const std::string& msg);
int32_t fail(int32_t code, const std::string& msg);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define AFB_CUDA(expr)                                                          \
    do {                                                                        \
        cudaError_t afb_e_ = (expr);                                            \
        if (afb_e_ != cudaSuccess) return afb::cuda_fail(afb_e_, #expr, __FILE__, __LINE__); \
    } while (0)
#define AFB_TRY(expr)                         \
    do {                                      \
        int32_t afb_s_ = (expr);              \
        if (afb_s_ != AFB_OK) return afb_s_;  \
    } while (0)
#define AFB_KERNEL_CHECK() AFB_CUDA(cudaGetLastError())

constexpr int MAX_DEVICES = 16;  // device ordinals 0..15 (one node); also the limit of the peer-pointer tables
int32_t ensure_device();  // AFB_NO_DEVICE unless a CUDA device is usable (no CPU fallback)
int current_device();     // the calling thread's current CUDA device ordinal
int sm_count();           // ... and its SM count

// Opt-in dynamic shared memory: the attribute belongs to the (kernel, device) pair, so it is set once per device.
#ifndef AFB_HOST_EMU
#define AFB_SMEM_OPTIN(kernel, bytes)                                                                          \
    do {                                                                                                       \
        static unsigned afb_done_mask_ = 0;  /* benign race: the attribute call is idempotent */                \
        const unsigned afb_bit_ = 1u << afb::current_device();                                                 \
        if (!(__atomic_load_n(&afb_done_mask_, __ATOMIC_ACQUIRE) & afb_bit_)) {                                 \
            AFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            __atomic_fetch_or(&afb_done_mask_, afb_bit_, __ATOMIC_RELEASE);                                    \
        }                                                                                                      \
    } while (0)
#else
#define AFB_SMEM_OPTIN(kernel, bytes) do { (void)(bytes); } while (0)
#endif

// ---- complex helpers ----------------------------------------------------------------------------
typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cmul_mi(cplx a) { return make_double2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ cplx cmul_pi(cplx a) { return make_double2(-a.y, a.x); }  // a * (+i)
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }

__device__ __forceinline__ cplx ldg2(const cplx* p) {
#ifdef AFB_HOST_EMU
    return *p;
#else
    return __ldg(reinterpret_cast<const double2*>(p));
#endif
}

// asynchronous global -> shared copies (LDGSTS): no register staging, many copies in flight per thread
__device__ __forceinline__ void async_copy8(double* smem_dst, const double* gmem_src) {
#ifdef AFB_HOST_EMU
    *smem_dst = *gmem_src;
#else
    __pipeline_memcpy_async(smem_dst, gmem_src, 8);
#endif
}
__device__ __forceinline__ void async_copy16(double* smem_dst, const double* gmem_src) {  // both 16-byte aligned
#ifdef AFB_HOST_EMU
    smem_dst[0] = gmem_src[0];
    smem_dst[1] = gmem_src[1];
#else
    __pipeline_memcpy_async(smem_dst, gmem_src, 16);
#endif
}
__device__ __forceinline__ void async_copy_commit() {
#ifndef AFB_HOST_EMU
    __pipeline_commit();
#endif
}
__device__ __forceinline__ void async_copy_wait() {
#ifndef AFB_HOST_EMU
    __pipeline_commit();
    __pipeline_wait_prior(0);
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* gmem) {
#ifndef AFB_HOST_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem));
#else
    (void)gmem;
#endif
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (sm_90+: SASS UBLKCP / SYNCS) ------------------
// One thread issues; every thread of the CTA waits on the barrier's phase bit.  bytes, dst and src multiples of 16.
__device__ __forceinline__ void mbar_init(uint64_t* bar) {
#ifdef AFB_HOST_EMU
    *bar = 0;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
#ifdef AFB_HOST_EMU
    std::memcpy(smem_dst, gmem_src, bytes);
    (void)bar;
#else
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    // the buffer was last read through the generic proxy: order those reads before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
#endif
}
// TMA 1-D bulk copy shared -> global (also peer-mapped global memory over NVLink), tracked by bulk async-groups
__device__ __forceinline__ void bulk_copy_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
#ifdef AFB_HOST_EMU
    std::memcpy(gmem_dst, smem_src, bytes);
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit_group() {
#ifndef AFB_HOST_EMU
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
// waits until at most N of the thread's bulk groups are still READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
#ifndef AFB_HOST_EMU
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
#endif
}
// ... until at most N have not completed (writes performed)
template <int N> __device__ __forceinline__ void bulk_wait_group() {
#ifndef AFB_HOST_EMU
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async_smem() {
#ifndef AFB_HOST_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
#ifndef AFB_HOST_EMU
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
#else
    (void)gmem_src; (void)bytes;
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
#ifndef AFB_HOST_EMU
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "AFB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra AFB_MBAR_DONE;\n"
        "bra AFB_MBAR_WAIT;\n"
        "AFB_MBAR_DONE:\n"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
#else
    (void)bar; (void)phase;
#endif
}

static inline int ilog2(int64_t v) {
    int l = 0;
    while ((int64_t(1) << (l + 1)) <= v) ++l;
    return l;
}
static inline bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

// ---- twiddle tables (device resident, built once per size, double-double accurate on host) ------
// root_table(N)[m] = exp(-2 pi i m / N), m = 0..N-1
int32_t root_table(int64_t N, const cplx** out);
// half_root_table(n)[k] = exp(-i pi k / (2 n)), k = 0..n   (DCT post/pre twiddles)
int32_t quarter_table(int64_t n, const cplx** out);
void free_tables();

}  // namespace afb
