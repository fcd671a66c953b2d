// FP64 Clenshaw evaluation, Clenshaw bisection (inverse-CDF sampling) and the normalised-cumsum
// helpers: kernels K6, K7, K9 of SURVEY.md section 2c.
//
// Arithmetic is written with explicit __fma_rn/__dsub_rn so that every result is bit-identical to the
// reference recurrence evaluated with a fused muladd (UPSTREAM AFOP clenshaw; call sites
// /root/reference/src/Extras/sample.jl:47,56,68,80,94):
//     x2 = 2x; b1 = b2 = 0; for k = N:-1:2: (b2,b1) = (b1, muladd(x2,b1,c[k]-b2)); muladd(x2/2,b1,c[1]-b2)
//
// Roofline: FP64 pipe.  One coefficient step costs one DFMA + one DADD per point (3 flops in two
// issue slots), so the structural ceiling is 75 % of the 2-flop/slot FMA peak.  Coefficients live in
// shared memory and are read as broadcast 128-bit loads (two coefficients per LDS); every thread
// carries R independent points so the DFMA dependency chains overlap.
#include <algorithm>

#include "afb_internal.h"
#ifndef AFB_HOST_EMU
#include <cuda_pipeline.h>
#endif

namespace afb {

constexpr int CL_THREADS = 512;        // threads per CTA
#ifndef AFB_CL_R
#define AFB_CL_R 4
#endif
constexpr int CL_R = AFB_CL_R;         // points per thread (independent recurrences) of the bisection / 2-D kernels
// ... and of the evaluation kernel: more points share every broadcast coefficient load and its .reuse'd DADD operand
// (measured N = 10^4, 2.7e8 points: 4 -> 21.9, 6 -> 23.35, 8 -> 22.9, 12 -> 21.9, 16 -> 22.2 TFLOP/s; 6 is the largest count
//  without spills at 64 registers.  The bisection kernel got slower with 8: 21.3 -> 19.8, it stays at 4.)
#ifndef AFB_CLV_R
#define AFB_CLV_R 6
#endif
constexpr int CLV_R = AFB_CLV_R;
constexpr int CL_CHUNK = 12288;        // coefficients staged per shared-memory chunk (96 KB)

// Runs the recurrence over coefficients c[klo..khi) (khi exclusive), from high to low, with klo >= 1.
// `c` may point to shared or global memory.  (b1, b2) follow the reference naming.
template <int R>
__device__ __forceinline__ void clenshaw_span(const double* c, int klo, int khi, const double (&x2)[R],
                                              double (&b1)[R], double (&b2)[R]) {
    int k = khi - 1;
    // peel to make (k - klo + 1) even, then two steps per iteration without register moves
    if (((k - klo + 1) & 1) && k >= klo) {
        const double ck = c[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double t = __fma_rn(x2[r], b1[r], __dsub_rn(ck, b2[r]));
            b2[r] = b1[r];
            b1[r] = t;
        }
        --k;
    }
    for (; k > klo; k -= 2) {
        const double ck = c[k], ck1 = c[k - 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            b2[r] = __fma_rn(x2[r], b1[r], __dsub_rn(ck, b2[r]));   // new b1 lives in b2
            b1[r] = __fma_rn(x2[r], b2[r], __dsub_rn(ck1, b1[r]));  // next step, roles swapped back
        }
    }
}

template <int R>
__device__ __forceinline__ void clenshaw_finish(double c0, const double (&x2)[R], const double (&b1)[R],
                                                const double (&b2)[R], double (&y)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) y[r] = __fma_rn(__dmul_rn(x2[r], 0.5), b1[r], __dsub_rn(c0, b2[r]));
}

// ------------------------------------------------------------------------------------------------
// K6: y[i] = sum_k c[k] T_k(x[i]),  one CTA loops over tiles of CL_THREADS*R points
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, CLV_R <= 8 ? 2 : 1)
clenshaw_vec_kernel(const double* __restrict__ c, int64_t N, const double* __restrict__ x,
                    double* __restrict__ y, int64_t P, int64_t ntiles) {
    AFB_DYN_SMEM(double, sc);
    const int tid = threadIdx.x;
    const int64_t nchunks = (N + CL_CHUNK - 1) / CL_CHUNK;
    if (nchunks <= 1) {
        for (int64_t k = tid; k < N; k += CL_THREADS) sc[k] = c[k];
        __syncthreads();
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * (int64_t)(CL_THREADS * CLV_R);
        double x2[CLV_R], b1[CLV_R], b2[CLV_R], out[CLV_R];
#pragma unroll
        for (int r = 0; r < CLV_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * CL_THREADS;
            const double xv = (i < P) ? x[i] : 0.0;
            x2[r] = __dmul_rn(2.0, xv);
            b1[r] = 0.0;
            b2[r] = 0.0;
        }
        if (N == 0) {
#pragma unroll
            for (int r = 0; r < CLV_R; ++r) out[r] = 0.0;
        } else if (nchunks <= 1) {
            clenshaw_span<CLV_R>(sc, 1, (int)N, x2, b1, b2);
            clenshaw_finish<CLV_R>(sc[0], x2, b1, b2, out);
        } else {
            double c0 = 0.0;
            for (int64_t ch = nchunks - 1; ch >= 0; --ch) {
                const int64_t lo = ch * CL_CHUNK;
                const int64_t hi = (lo + CL_CHUNK < N) ? lo + CL_CHUNK : N;
                __syncthreads();  // previous chunk fully consumed
                for (int64_t k = lo + tid; k < hi; k += CL_THREADS) sc[k - lo] = c[k];
                __syncthreads();
                const int klo = (ch == 0) ? 1 : 0;
                clenshaw_span<CLV_R>(sc, klo, (int)(hi - lo), x2, b1, b2);
                if (ch == 0) c0 = sc[0];
            }
            clenshaw_finish<CLV_R>(c0, x2, b1, b2, out);
        }
#pragma unroll
        for (int r = 0; r < CLV_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * CL_THREADS;
            if (i < P) y[i] = (N == 1) ? c[0] : out[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011); 53-bit uniforms in [0,1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&ctr)[4], const uint32_t (&key)[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ ctr[1] ^ key[0];
    const uint32_t n2 = hi0 ^ ctr[3] ^ key[1];
    ctr[0] = n0; ctr[1] = lo1; ctr[2] = n2; ctr[3] = lo0;
}
// two uniforms per 128-bit block; index i uses block i>>1, half i&1
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t index) {
    const uint64_t blk = index >> 1;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(ctr, key);
        key[0] += 0x9E3779B9u;
        key[1] += 0xBB67AE85u;
    }
    const uint64_t w = (index & 1) ? (((uint64_t)ctr[3] << 32) | ctr[2]) : (((uint64_t)ctr[1] << 32) | ctr[0]);
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void philox_uniform_kernel(uint64_t seed, uint64_t offset, double* __restrict__ out, int64_t P) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = philox_uniform(seed, offset + (uint64_t)i);
}

// ------------------------------------------------------------------------------------------------
// K7: Clenshaw bisection (src/Extras/sample.jl:58-75), fused with fromcanonical (sample.jl:106)
//     a=-1,b=1; numits x { m=.5(a+b); v=clenshaw(c,m); (v<=u) ? a=m : b=m }; out = .5(a+b)
// USE_RNG: uniforms drawn on device (throughput path of `sample`), else read from `u`.
// ------------------------------------------------------------------------------------------------
template <bool USE_RNG, bool SMEM_COEF>
__global__ void __launch_bounds__(CL_THREADS, 2)
bisect_vec_kernel(const double* __restrict__ c, int64_t N, const double* __restrict__ u, uint64_t seed,
                  uint64_t offset, double* __restrict__ out, int64_t P, int numits, double da, double db,
                  int64_t ntiles) {
    AFB_DYN_SMEM(double, sc);
    const int tid = threadIdx.x;
    const double* cc = c;
    if (SMEM_COEF) {
        for (int64_t k = tid; k < N; k += CL_THREADS) sc[k] = c[k];
        __syncthreads();
        cc = sc;
    }
    const double mid = __dmul_rn(__dadd_rn(da, db), 0.5);  // mean(d)
    const double len = __dsub_rn(db, da);                   // complexlength(d)
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * (int64_t)(CL_THREADS * CL_R);
        double a[CL_R], b[CL_R], uu[CL_R];
#pragma unroll
        for (int r = 0; r < CL_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * CL_THREADS;
            a[r] = -1.0;
            b[r] = 1.0;
            if (USE_RNG) uu[r] = philox_uniform(seed, offset + (uint64_t)i);
            else uu[r] = (i < P) ? u[i] : 0.0;
        }
        for (int it = 0; it < numits; ++it) {
            double m[CL_R], x2[CL_R], b1[CL_R], b2[CL_R], v[CL_R];
#pragma unroll
            for (int r = 0; r < CL_R; ++r) {
                m[r] = __dmul_rn(0.5, __dadd_rn(a[r], b[r]));
                x2[r] = __dmul_rn(2.0, m[r]);
                b1[r] = 0.0;
                b2[r] = 0.0;
            }
            if (N == 0) {
#pragma unroll
                for (int r = 0; r < CL_R; ++r) v[r] = 0.0;
            } else if (N == 1) {
#pragma unroll
                for (int r = 0; r < CL_R; ++r) v[r] = cc[0];
            } else {
                clenshaw_span<CL_R>(cc, 1, (int)N, x2, b1, b2);
                clenshaw_finish<CL_R>(cc[0], x2, b1, b2, v);
            }
#pragma unroll
            for (int r = 0; r < CL_R; ++r) {
                if (v[r] <= uu[r]) a[r] = m[r]; else b[r] = m[r];
            }
        }
#pragma unroll
        for (int r = 0; r < CL_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * CL_THREADS;
            const double xm = __dmul_rn(0.5, __dadd_rn(a[r], b[r]));
            // fromcanonical(d, x) = mean(d) + complexlength(d) x / 2
            if (i < P) out[i] = __dadd_rn(mid, __dmul_rn(__dmul_rn(len, xm), 0.5));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Matrix-coefficient variants (column j of C at x[j]); src/Extras/sample.jl:80-101.
// A CTA owns W consecutive columns; when the N x W tile fits in shared memory it is staged there
// transposed (sm[k*W + col]) with coalesced global reads, else columns are walked in global memory.
// ------------------------------------------------------------------------------------------------
constexpr int COLS_W = 128;
constexpr int COLS_LD = COLS_W + 1;  // padded row of the transposed tile (conflict-free both ways)

template <bool BISECT>
__global__ void __launch_bounds__(COLS_W)
cols_kernel(const double* __restrict__ C, int64_t N, int64_t ld, const double* __restrict__ xu,
            double* __restrict__ out, int64_t P, int numits, int use_smem) {
    AFB_DYN_SMEM(double, sm);
    const int tid = threadIdx.x;
    const int64_t col0 = (int64_t)blockIdx.x * COLS_W;
    const int64_t j = col0 + tid;
    if (use_smem) {
        // coalesced: the tile is (nearly) contiguous in global memory; consecutive threads read consecutive k,
        // and the padded row stride makes the transposed shared-memory writes conflict free
        const int64_t ncol = (P - col0 < COLS_W) ? (P - col0) : COLS_W;
        const double* src = C + col0 * ld;
        const int n32 = (int)N, total = (int)ncol * n32;
        int cidx = tid / n32, k = tid - cidx * n32;          // one division, then incremental (cidx, k) updates
        for (int e = tid; e < total; e += COLS_W) {
            async_copy8(sm + k * COLS_LD + cidx, src + (int64_t)cidx * ld + k);
            k += COLS_W;
            while (k >= n32) { k -= n32; ++cidx; }
        }
        async_copy_wait();
        __syncthreads();
    }
    if (j >= P) return;
    const double* gcol = C + j * ld;
    auto coef = [&](int64_t k) -> double { return use_smem ? sm[k * COLS_LD + tid] : gcol[k]; };
    auto eval = [&](double x) -> double {
        if (N == 0) return 0.0;
        if (N == 1) return coef(0);
        const double x2 = __dmul_rn(2.0, x);
        double b1 = 0.0, b2 = 0.0;
        for (int64_t k = N - 1; k >= 1; --k) {
            const double t = __fma_rn(x2, b1, __dsub_rn(coef(k), b2));
            b2 = b1;
            b1 = t;
        }
        return __fma_rn(__dmul_rn(x2, 0.5), b1, __dsub_rn(coef(0), b2));
    };
    if (!BISECT) {
        out[j] = eval(xu[j]);
    } else {
        double a = -1.0, b = 1.0;
        const double u = xu[j];
        for (int it = 0; it < numits; ++it) {
            const double m = __dmul_rn(0.5, __dadd_rn(a, b));
            if (eval(m) <= u) a = m; else b = m;
        }
        out[j] = __dmul_rn(0.5, __dadd_rn(a, b));
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-function evaluation  Y[m + M*i] = f_m(x[i])   (evaluate.(f.A, ry'), sample.jl:208)
// Coefficients (N x M) are staged in shared memory when they fit; all lanes read the same
// coefficient (broadcast), each thread owns one point and walks the M functions.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
multi_kernel(const double* __restrict__ C, int64_t N, int64_t M, const double* __restrict__ x,
             double* __restrict__ Y, int64_t P, int use_smem) {
    AFB_DYN_SMEM(double, sm);
    const int tid = threadIdx.x;
    const double* cc = C;
    if (use_smem) {
        for (int64_t k = tid; k < N * M; k += blockDim.x) sm[k] = C[k];
        __syncthreads();
        cc = sm;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const double xv = x[i];
        const double x2 = __dmul_rn(2.0, xv);
        for (int64_t m = 0; m < M; ++m) {
            const double* c = cc + m * N;
            double r;
            if (N == 0) r = 0.0;
            else if (N == 1) r = c[0];
            else {
                double b1 = 0.0, b2 = 0.0;
                for (int64_t k = N - 1; k >= 1; --k) {
                    const double t = __fma_rn(x2, b1, __dsub_rn(c[k], b2));
                    b2 = b1;
                    b1 = t;
                }
                r = __fma_rn(__dmul_rn(x2, 0.5), b1, __dsub_rn(c[0], b2));
            }
            Y[m + M * i] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bivariate Clenshaw: f(x_p, y_p) = sum_{i,j} C[i,j] T_i(x_p) T_j(y_p), C column-major n1 x n2.
// Evaluation order of a ProductFun (UPSTREAM ApproxFunBase): every column function g_j(x) = sum_i C[i,j] T_i(x)
// by the 1-D recurrence, then the 1-D recurrence over j in y with the g_j as coefficients -- fused here so that
// the n2 x P intermediate never exists.  Coefficients in shared memory (broadcast reads), R points per thread.
// ------------------------------------------------------------------------------------------------
constexpr int C2_THREADS = 256;
__global__ void __launch_bounds__(C2_THREADS)
clenshaw2d_kernel(const double* __restrict__ C, int64_t n1, int64_t n2, const double* __restrict__ x,
                  const double* __restrict__ y, double* __restrict__ out, int64_t P, int64_t ntiles, int use_smem) {
    AFB_DYN_SMEM(double, sc);
    const int tid = threadIdx.x;
    const double* cc = C;
    if (use_smem) {
        for (int64_t k = tid; k < n1 * n2; k += C2_THREADS) sc[k] = C[k];
        __syncthreads();
        cc = sc;
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * (int64_t)(C2_THREADS * CL_R);
        double x2[CL_R], y2[CL_R], B1[CL_R], B2[CL_R], res[CL_R];
#pragma unroll
        for (int r = 0; r < CL_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * C2_THREADS;
            x2[r] = __dmul_rn(2.0, (i < P) ? x[i] : 0.0);
            y2[r] = __dmul_rn(2.0, (i < P) ? y[i] : 0.0);
            B1[r] = 0.0; B2[r] = 0.0; res[r] = 0.0;
        }
        for (int64_t j = n2 - 1; j >= 0; --j) {
            const double* cj = cc + j * n1;
            double g[CL_R];
            if (n1 == 0) {
#pragma unroll
                for (int r = 0; r < CL_R; ++r) g[r] = 0.0;
            } else if (n1 == 1) {
#pragma unroll
                for (int r = 0; r < CL_R; ++r) g[r] = cj[0];
            } else {
                double b1[CL_R], b2[CL_R];
#pragma unroll
                for (int r = 0; r < CL_R; ++r) { b1[r] = 0.0; b2[r] = 0.0; }
                clenshaw_span<CL_R>(cj, 1, (int)n1, x2, b1, b2);
                clenshaw_finish<CL_R>(cj[0], x2, b1, b2, g);
            }
            if (j >= 1) {
#pragma unroll
                for (int r = 0; r < CL_R; ++r) {
                    const double t = __fma_rn(y2[r], B1[r], __dsub_rn(g[r], B2[r]));
                    B2[r] = B1[r];
                    B1[r] = t;
                }
            } else {
#pragma unroll
                for (int r = 0; r < CL_R; ++r)
                    res[r] = (n2 == 1) ? g[r] : __fma_rn(__dmul_rn(y2[r], 0.5), B1[r], __dsub_rn(g[r], B2[r]));
            }
        }
#pragma unroll
        for (int r = 0; r < CL_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * C2_THREADS;
            if (i < P) out[i] = res[r];
        }
    }
}

int32_t launch_clenshaw2d(const double* d_C, int64_t n1, int64_t n2, const double* d_x, const double* d_y, double* d_out,
                          int64_t P, void* stream) {
    if (P == 0) return AFB_OK;
    const int64_t ntiles = (P + (int64_t)C2_THREADS * CL_R - 1) / ((int64_t)C2_THREADS * CL_R);
    const int64_t grid = ntiles < 4 * (int64_t)sm_count() ? ntiles : 4 * (int64_t)sm_count();
    const size_t need = sizeof(double) * (size_t)(n1 * n2);
    const int use_smem = (n1 * n2 > 0 && need <= 96 * 1024) ? 1 : 0;
    auto k = clenshaw2d_kernel;
    AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    AFB_LAUNCH(k, (unsigned)grid, C2_THREADS, use_smem ? need : 8, stream, d_C, n1, n2, d_x, d_y, d_out, P, ntiles, use_smem);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Fourier (x) Chebyshev evaluation: column j of C (length n1, Fourier order [a0, b1, a1, b2, a2, ...]) is the trigonometric
// series g_j(t) = a0 + sum_k (b_k sin kt + a_k cos kt), summed by the two Clenshaw recurrences in c = cos t
// (sum a_k T_k(c)  and  sin t * sum b_k U_{k-1}(c)); the g_j are then the coefficients of the Chebyshev recurrence in y.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fourier_cheb2d_kernel(const double* __restrict__ C, int64_t n1, int64_t n2, const double* __restrict__ t,
                      const double* __restrict__ y, double* __restrict__ out, int64_t P, int use_smem) {
    AFB_DYN_SMEM(double, sc);
    const double* cc = C;
    if (use_smem) {
        for (int64_t k = threadIdx.x; k < n1 * n2; k += blockDim.x) sc[k] = C[k];
        __syncthreads();
        cc = sc;
    }
    const int64_t K = n1 / 2;          // highest frequency with a sine slot (slot 2K-1)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        double sn, cs;
        sincos(t[i], &sn, &cs);
        const double c2 = __dmul_rn(2.0, cs), y2 = __dmul_rn(2.0, y[i]);
        double B1 = 0.0, B2 = 0.0, res = 0.0;
        for (int64_t j = n2 - 1; j >= 0; --j) {
            const double* cj = cc + j * n1;
            double g = 0.0;
            if (n1 >= 1) {
                // cosine part: a_k = cj[2k], k = 0..(n1-1)/2
                double b1 = 0.0, b2 = 0.0;
                for (int64_t k = (n1 - 1) / 2; k >= 1; --k) {
                    const double tt = __fma_rn(c2, b1, __dsub_rn(cj[2 * k], b2));
                    b2 = b1; b1 = tt;
                }
                g = __fma_rn(cs, b1, __dsub_rn(cj[0], b2));
                // sine part: b_k = cj[2k-1], k = 1..K:  d_k = b_k + 2c d_{k+1} - d_{k+2};  sum = sin t * d_1
                double d1 = 0.0, d2 = 0.0;
                for (int64_t k = K; k >= 1; --k) {
                    const double tt = __fma_rn(c2, d1, __dsub_rn(cj[2 * k - 1], d2));
                    d2 = d1; d1 = tt;
                }
                g = __fma_rn(sn, d1, g);
            }
            if (j >= 1) {
                const double tt = __fma_rn(y2, B1, __dsub_rn(g, B2));
                B2 = B1; B1 = tt;
            } else {
                res = (n2 == 1) ? g : __fma_rn(__dmul_rn(y2, 0.5), B1, __dsub_rn(g, B2));
            }
        }
        out[i] = res;
    }
}
int32_t launch_fourier_cheb2d(const double* d_C, int64_t n1, int64_t n2, const double* d_t, const double* d_y, double* d_out,
                              int64_t P, void* stream) {
    if (P == 0) return AFB_OK;
    int64_t grid = (P + 255) / 256;
    if (grid > 8 * (int64_t)sm_count()) grid = 8 * (int64_t)sm_count();
    const size_t need = sizeof(double) * (size_t)(n1 * n2);
    const int use_smem = (n1 * n2 > 0 && need <= 96 * 1024) ? 1 : 0;
    auto k = fourier_cheb2d_kernel;
    AFB_SMEM_OPTIN(k, 96 * 1024);
    AFB_LAUNCH(k, (unsigned)grid, 256, use_smem ? need : 8, stream, d_C, n1, n2, d_t, d_y, d_out, P, use_smem);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ------------------------------------------------------------------------------------------------
// K9: chebnormalizedcumsum! per column (sample.jl:170-175), sequential per column to keep the
// reference's summation order; one thread per column.
// ------------------------------------------------------------------------------------------------
__global__ void normcumsum_kernel(double* __restrict__ C, int64_t N, int64_t ncols, int64_t ld, int grow) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double* v = C + j * ld;
    const int64_t n = N;
    if (n == 2) {
        v[1] = __dmul_rn(v[1], 0.5);
    } else if (n > 2) {
        v[0] = __dsub_rn(v[0], __dmul_rn(v[2], 0.5));
        for (int64_t q = 1; q <= n - 3; ++q) v[q] = __dmul_rn(__dsub_rn(v[q], v[q + 2]), 0.5);
        v[n - 2] = __dmul_rn(v[n - 2], 0.5);
        v[n - 1] = __dmul_rn(v[n - 1], 0.5);
    }
    const int64_t L = grow ? n + 1 : n;
    for (int64_t k = L - 1; k >= 1; --k) v[k] = __ddiv_rn(v[k - 1], (double)k);
    if (L > 0) v[0] = 0.0;
    for (int64_t k = 2; k <= L; ++k) v[0] = __dadd_rn(v[0], (k & 1) ? -v[k - 1] : v[k - 1]);
    double val = 0.0;
    for (int64_t k = 0; k < L; ++k) val = __dadd_rn(val, v[k]);
    val = __ddiv_rn(1.0, val);
    for (int64_t k = 0; k < L; ++k) v[k] = __dmul_rn(v[k], val);
}

// Tiled variant: a CTA stages 128 columns transposed in shared memory (coalesced both ways) and each thread runs
// the same sequential column recipe on its shared-memory column.  Same operation order => same bits.
__global__ void __launch_bounds__(COLS_W)
normcumsum_tiled_kernel(double* __restrict__ C, int64_t N, int64_t ncols, int64_t ld, int grow) {
    AFB_DYN_SMEM(double, sm);
    const int tid = threadIdx.x;
    const int64_t col0 = (int64_t)blockIdx.x * COLS_W;
    const int64_t ncol = (ncols - col0 < COLS_W) ? (ncols - col0) : COLS_W;
    const int64_t n = N, L = grow ? N + 1 : N;
    {
        const double* src = C + col0 * ld;
        const int n32 = (int)n, total = (int)ncol * n32;
        int cidx = tid / n32, k = tid - cidx * n32;
        for (int e = tid; e < total; e += COLS_W) {
            async_copy8(sm + k * COLS_LD + cidx, src + (int64_t)cidx * ld + k);
            k += COLS_W;
            while (k >= n32) { k -= n32; ++cidx; }
        }
        async_copy_wait();
    }
    __syncthreads();
    if (tid < ncol) {
        double* v = sm + tid;
#define V(k) v[(k) * COLS_LD]
        if (n == 2) {
            V(1) = __dmul_rn(V(1), 0.5);
        } else if (n > 2) {
            V(0) = __dsub_rn(V(0), __dmul_rn(V(2), 0.5));
            for (int64_t q = 1; q <= n - 3; ++q) V(q) = __dmul_rn(__dsub_rn(V(q), V(q + 2)), 0.5);
            V(n - 2) = __dmul_rn(V(n - 2), 0.5);
            V(n - 1) = __dmul_rn(V(n - 1), 0.5);
        }
        for (int64_t k = L - 1; k >= 1; --k) V(k) = __ddiv_rn(V(k - 1), (double)k);
        if (L > 0) V(0) = 0.0;
        for (int64_t k = 2; k <= L; ++k) V(0) = __dadd_rn(V(0), (k & 1) ? -V(k - 1) : V(k - 1));
        double val = 0.0;
        for (int64_t k = 0; k < L; ++k) val = __dadd_rn(val, V(k));
        val = __ddiv_rn(1.0, val);
        for (int64_t k = 0; k < L; ++k) V(k) = __dmul_rn(V(k), val);
#undef V
    }
    __syncthreads();
    {
        double* dst = C + col0 * ld;
        const int l32 = (int)L, total = (int)ncol * l32;
        int cidx = tid / l32, k = tid - cidx * l32;
        for (int e = tid; e < total; e += COLS_W) {
            dst[(int64_t)cidx * ld + k] = sm[k * COLS_LD + cidx];
            k += COLS_W;
            while (k >= l32) { k -= l32; ++cidx; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// points(::Chebyshev, n), descending (NEWS.md:92): x_k = sinpi((n-2k-1)/(2n)) mapped to [a,b]
// ------------------------------------------------------------------------------------------------
__global__ void chebpoints_kernel(double* __restrict__ out, int64_t n, double da, double db) {
    const double mid = __dmul_rn(__dadd_rn(da, db), 0.5);
    const double len = __dsub_rn(db, da);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double arg = __ddiv_rn((double)(n - 2 * k - 1), (double)(2 * n));
        const double xc = sinpi(arg);
        out[k] = __dadd_rn(mid, __dmul_rn(__dmul_rn(len, xc), 0.5));
    }
}

// ================================================================================================
// launchers
// ================================================================================================
// ------------------------------------------------------------------------------------------------
// General three-term Clenshaw (Ultraspherical / Jacobi / Legendre evaluation):
//   p_{j+1} = (A_j x + B_j) p_j - C_j p_{j-1},  y = sum_j c_j p_j(x)
//   for j = N-1 .. 0:  bk2, bk1 = bk1, fma(fma(A_j, x, B_j), bk1, fma(-C_{j+1}, bk2, c_j))      y = bk1
// (the operation order of UPSTREAM `clenshaw(sp::PolynomialSpace, c, x)`; its last step is the same formula at j = 0).
// Degrees are staged in shared memory four doubles each {c_j, A_j, B_j, -C_{j+1}} (two 128-bit broadcast loads per
// step); every thread runs four independent points.
// ------------------------------------------------------------------------------------------------
constexpr int REC_THREADS = 256;
constexpr int REC_R = 4;
constexpr int REC_CHUNK = 3072;  // degrees per chunk: 96 KB
__global__ void __launch_bounds__(REC_THREADS, 2)
clenshaw_rec_kernel(const double* __restrict__ c, int64_t N, const double* __restrict__ A, const double* __restrict__ B,
                    const double* __restrict__ C, const double* __restrict__ x, double* __restrict__ y, int64_t P,
                    int64_t ntiles) {
    AFB_DYN_SMEM(double, sc);
    const int tid = threadIdx.x;
    const int64_t nchunks = (N + REC_CHUNK - 1) / REC_CHUNK;
    auto stage = [&](int64_t lo, int64_t hi) {
        for (int64_t j = lo + tid; j < hi; j += REC_THREADS) {
            double* q = sc + 4 * (j - lo);
            q[0] = c[j]; q[1] = A[j]; q[2] = B[j]; q[3] = -C[j + 1];
        }
    };
    if (nchunks == 1) {
        stage(0, N);
        __syncthreads();
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * (int64_t)(REC_THREADS * REC_R);
        double xv[REC_R], b1[REC_R], b2[REC_R];
#pragma unroll
        for (int r = 0; r < REC_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * REC_THREADS;
            xv[r] = (i < P) ? x[i] : 0.0;
            b1[r] = 0.0;
            b2[r] = 0.0;
        }
        for (int64_t ch = nchunks - 1; ch >= 0; --ch) {
            const int64_t lo = ch * REC_CHUNK;
            const int64_t hi = (lo + REC_CHUNK < N) ? lo + REC_CHUNK : N;
            if (nchunks > 1) {
                __syncthreads();  // previous chunk fully consumed
                stage(lo, hi);
                __syncthreads();
            }
            for (int k = (int)(hi - lo) - 1; k >= 0; --k) {
                const double2 ca = *reinterpret_cast<const double2*>(sc + 4 * k);
                const double2 bc = *reinterpret_cast<const double2*>(sc + 4 * k + 2);
#pragma unroll
                for (int r = 0; r < REC_R; ++r) {
                    const double t = __fma_rn(__fma_rn(ca.y, xv[r], bc.x), b1[r], __fma_rn(bc.y, b2[r], ca.x));
                    b2[r] = b1[r];
                    b1[r] = t;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < REC_R; ++r) {
            const int64_t i = base + tid + (int64_t)r * REC_THREADS;
            if (i < P) y[i] = b1[r];  // N == 0: no step ran, y = 0
        }
    }
}

int32_t launch_clenshaw_rec(const double* d_c, int64_t N, const double* d_A, const double* d_B, const double* d_C,
                            const double* d_x, double* d_y, int64_t P, void* stream) {
    if (P == 0) return AFB_OK;
    const int64_t per = (int64_t)REC_THREADS * REC_R;
    const int64_t ntiles = (P + per - 1) / per;
    const int64_t grid = ntiles < 2 * (int64_t)sm_count() ? ntiles : 2 * (int64_t)sm_count();
    const int64_t staged = N < REC_CHUNK ? (N > 0 ? N : 1) : REC_CHUNK;
    auto k = clenshaw_rec_kernel;
    AFB_SMEM_OPTIN(k, (sizeof(double) * 4 * REC_CHUNK));
    AFB_LAUNCH(k, (unsigned)grid, REC_THREADS, sizeof(double) * 4 * (size_t)staged, stream, d_c, N, d_A, d_B, d_C, d_x, d_y, P,
               ntiles);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Piecewise Chebyshev Funs (PiecewiseSpace of Chebyshev segments): evaluation and the GENERIC bisection
// of src/Extras/sample.jl:8-36 (the path `sample(abs(Fun(sin, -5..5)), n)` takes, test/PiecewiseSampleTest.jl:5-10):
//     a = minimum(d); b = maximum(d); numits x { m = genmidpoint(a,b); v = f(m); (v <= u) ? a = m : b = m }; .5(a+b)
// f(m) = the FIRST piece whose closed interval contains m, evaluated by the Chebyshev recurrence at
// tocanonical(piece, m) = (2m - lo - hi)/(hi - lo); zero outside the domain.
// Coefficients of all pieces (concatenated, offs[k]..offs[k+1]) and the np+1 break points live in shared memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double genmidpoint_dev(double a, double b) {  // sample.jl:11-21
    const bool ia = isinf(a), ib = isinf(b);
    if (ia && ib) return 0.0;
    if (ia) return __dsub_rn(b, 100.0);
    if (ib) return __dadd_rn(a, 100.0);
    return __dmul_rn(__dadd_rn(a, b), 0.5);
}
__device__ __forceinline__ double piecewise_eval(const double* sc, const int32_t* soff, const double* sbrk, int np, double m) {
    if (!(m >= sbrk[0] && m <= sbrk[np])) return 0.0;
    int lo = 0, hi = np - 1;  // smallest k with m <= brk[k+1]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (m <= sbrk[mid + 1]) hi = mid; else lo = mid + 1;
    }
    const double l = sbrk[lo], h = sbrk[lo + 1];
    const int o = soff[lo], N = soff[lo + 1] - o;
    if (N == 0) return 0.0;
    if (N == 1) return sc[o];
    const double x = __ddiv_rn(__dsub_rn(__dsub_rn(__dmul_rn(2.0, m), l), h), __dsub_rn(h, l));
    double x2[1] = {__dmul_rn(2.0, x)}, b1[1] = {0.0}, b2[1] = {0.0}, y[1];
    clenshaw_span<1>(sc + o, 1, N, x2, b1, b2);
    clenshaw_finish<1>(sc[o], x2, b1, b2, y);
    return y[0];
}
__global__ void __launch_bounds__(256)
piecewise_kernel(const double* __restrict__ coef, const int32_t* __restrict__ offs, const double* __restrict__ brk, int np,
                 const double* __restrict__ xu, double* __restrict__ out, int64_t P, int bisect, int numits) {
    AFB_DYN_SMEM(double, sc);
    const int total = offs[np];
    double* sbrk = sc + total;
    int32_t* soff = reinterpret_cast<int32_t*>(sbrk + np + 1);
    for (int k = threadIdx.x; k < total; k += blockDim.x) sc[k] = coef[k];
    for (int k = threadIdx.x; k <= np; k += blockDim.x) { sbrk[k] = brk[k]; soff[k] = offs[k]; }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = xu[i];
        if (!bisect) { out[i] = piecewise_eval(sc, soff, sbrk, np, v); continue; }
        double a = sbrk[0], b = sbrk[np];
        for (int it = 0; it < numits; ++it) {
            const double m = genmidpoint_dev(a, b);
            if (piecewise_eval(sc, soff, sbrk, np, m) <= v) a = m; else b = m;
        }
        out[i] = __dmul_rn(0.5, __dadd_rn(a, b));
    }
}
int32_t launch_piecewise(const double* d_coef, const int32_t* d_offs, const double* d_brk, int64_t np, int64_t total,
                         const double* d_xu, double* d_out, int64_t P, bool bisect, int numits, void* stream) {
    if (P == 0) return AFB_OK;
    const size_t smem = sizeof(double) * (size_t)(total + np + 1) + sizeof(int32_t) * (size_t)(np + 1) + 16;
    if (smem > 200 * 1024) return fail(AFB_UNSUPPORTED, "piecewise Fun: more than 200 KB of coefficients and break points");
    int64_t grid = (P + 255) / 256;
    if (grid > 8 * (int64_t)sm_count()) grid = 8 * (int64_t)sm_count();
    auto k = piecewise_kernel;
    AFB_SMEM_OPTIN(k, 200 * 1024);
    AFB_LAUNCH(k, (unsigned)grid, 256, smem, stream, d_coef, d_offs, d_brk, (int)np, d_xu, d_out, P, bisect ? 1 : 0, numits);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

static inline int64_t cl_tiles(int64_t P) { return (P + (int64_t)CL_THREADS * CL_R - 1) / ((int64_t)CL_THREADS * CL_R); }

int32_t launch_clenshaw_vec(const double* d_c, int64_t N, const double* d_x, double* d_y, int64_t P, void* stream) {
    if (P == 0) return AFB_OK;
    const int64_t ntiles = (P + (int64_t)CL_THREADS * CLV_R - 1) / ((int64_t)CL_THREADS * CLV_R);
    const int64_t grid = ntiles < 2 * (int64_t)sm_count() ? ntiles : 2 * (int64_t)sm_count();
    const size_t smem = sizeof(double) * (size_t)((N < CL_CHUNK ? (N > 0 ? N : 1) : CL_CHUNK));
    auto k = clenshaw_vec_kernel;
    AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CL_CHUNK)));
    AFB_LAUNCH(k, (unsigned)grid, CL_THREADS, smem, stream, d_c, N, d_x, d_y, P, ntiles);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_bisect_vec(const double* d_c, int64_t N, const double* d_u, bool rng, uint64_t seed, uint64_t offset,
                          double* d_out, int64_t P, int numits, double a, double b, void* stream) {
    if (P == 0) return AFB_OK;
    const int64_t ntiles = cl_tiles(P);
    const int64_t grid = ntiles < 2 * (int64_t)sm_count() ? ntiles : 2 * (int64_t)sm_count();
    const bool in_smem = N <= CL_CHUNK;
    const size_t smem = in_smem ? sizeof(double) * (size_t)(N > 0 ? N : 1) : 8;
#define AFB_BISECT_LAUNCH(RNG, SM)                                                                          \
    do {                                                                                                    \
        auto k = bisect_vec_kernel<RNG, SM>;                                                                \
        AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                      (int)(sizeof(double) * CL_CHUNK)));                                   \
        AFB_LAUNCH(k, (unsigned)grid, CL_THREADS, smem, stream, d_c, N, d_u, seed, offset, d_out, P, numits, \
                   a, b, ntiles);                                                                           \
    } while (0)
    if (rng) { if (in_smem) AFB_BISECT_LAUNCH(true, true); else AFB_BISECT_LAUNCH(true, false); }
    else     { if (in_smem) AFB_BISECT_LAUNCH(false, true); else AFB_BISECT_LAUNCH(false, false); }
#undef AFB_BISECT_LAUNCH
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_philox(uint64_t seed, uint64_t offset, double* d_out, int64_t P, void* stream) {
    if (P == 0) return AFB_OK;
    int64_t grid = (P + 255) / 256;
    if (grid > 8 * (int64_t)sm_count()) grid = 8 * (int64_t)sm_count();
    auto k = philox_uniform_kernel;
    AFB_LAUNCH(k, (unsigned)grid, 256, 0, stream, seed, offset, d_out, P);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_cols(bool bisect, const double* d_C, int64_t N, int64_t ld, const double* d_xu, double* d_out,
                    int64_t P, int numits, void* stream) {
    if (P == 0) return AFB_OK;
    const int64_t grid = (P + COLS_W - 1) / COLS_W;
    const size_t need = sizeof(double) * (size_t)N * COLS_LD;
    const int use_smem = (N > 0 && need <= 160 * 1024) ? 1 : 0;
    const size_t smem = use_smem ? need : 8;
    if (bisect) {
        auto k = cols_kernel<true>;
        AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        AFB_LAUNCH(k, (unsigned)grid, COLS_W, smem, stream, d_C, N, ld, d_xu, d_out, P, numits, use_smem);
    } else {
        auto k = cols_kernel<false>;
        AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        AFB_LAUNCH(k, (unsigned)grid, COLS_W, smem, stream, d_C, N, ld, d_xu, d_out, P, numits, use_smem);
    }
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_multi(const double* d_C, int64_t N, int64_t M, const double* d_x, double* d_Y, int64_t P, void* stream) {
    if (P == 0 || M == 0) return AFB_OK;
    const size_t need = sizeof(double) * (size_t)(N * M);
    const int use_smem = (N * M > 0 && need <= 160 * 1024) ? 1 : 0;
    int64_t grid = (P + 255) / 256;
    if (grid > 8 * (int64_t)sm_count()) grid = 8 * (int64_t)sm_count();
    auto k = multi_kernel;
    AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    AFB_LAUNCH(k, (unsigned)grid, 256, use_smem ? need : 8, stream, d_C, N, M, d_x, d_Y, P, use_smem);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_normcumsum(double* d_C, int64_t N, int64_t ncols, int64_t ld, int grow, void* stream) {
    if (ncols == 0) return AFB_OK;
    const size_t need = sizeof(double) * (size_t)(N + 1) * COLS_LD;
    if (ncols >= 64 && need <= 200 * 1024) {
        auto kt = normcumsum_tiled_kernel;
        AFB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        AFB_LAUNCH(kt, (unsigned)((ncols + COLS_W - 1) / COLS_W), COLS_W, need, stream, d_C, N, ncols, ld, grow);
        AFB_KERNEL_CHECK();
        return AFB_OK;
    }
    auto k = normcumsum_kernel;
    AFB_LAUNCH(k, (unsigned)((ncols + 127) / 128), 128, 0, stream, d_C, N, ncols, ld, grow);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_chebpoints(double* d_out, int64_t n, double a, double b, void* stream) {
    if (n == 0) return AFB_OK;
    int64_t grid = (n + 255) / 256;
    if (grid > 4 * (int64_t)sm_count()) grid = 4 * (int64_t)sm_count();
    auto k = chebpoints_kernel;
    AFB_LAUNCH(k, (unsigned)grid, 256, 0, stream, d_out, n, a, b);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused conditional sampler of sample(f::LowRankFun{Chebyshev,Chebyshev}, n)  (src/Extras/sample.jl:205-214):
//     fA = evaluate.(f.A, ry') ; AB = CB*fA ; chebnormalizedcumsum!(AB) ; rx = chebbisectioninv(AB, rand(n))
// One thread per sample: its column of AB (N coefficients) lives in a padded shared-memory tile and never
// touches HBM -- the reference materialises the N x n matrix AB (and streams it 47 times).
//   A  : NA x R column-major (coefficients of the R functions f.A), evaluated at tocanonical(ya..yb, ry)
//   CB : N  x R column-major (coefficientmatrix(f.B))
// AB[k] = sum_r CB[k,r] fA[r] is accumulated in r order with fused multiply-adds.
// ------------------------------------------------------------------------------------------------
constexpr int S2_W = 128;
constexpr int S2_LD = S2_W + 1;
__global__ void __launch_bounds__(S2_W)
sample2d_kernel(const double* __restrict__ A, int64_t NA, const double* __restrict__ CB, int64_t N, int64_t R,
                double ya, double yb, double xa, double xb, const double* __restrict__ ry,
                const double* __restrict__ u, double* __restrict__ rx, int64_t P, int numits) {
    AFB_DYN_SMEM(double, sm);  // [N][S2_LD]
    const int tid = threadIdx.x;
    const int64_t j = (int64_t)blockIdx.x * S2_W + tid;
    const bool act = j < P;
    double* col = sm + tid;
    // tocanonical(d, y) = (2y - a - b) / (b - a); the reference's evaluate returns 0 outside the domain
    const double y = act ? ry[j] : 0.5 * (ya + yb);
    const bool inside = (y >= (ya < yb ? ya : yb)) && (y <= (ya < yb ? yb : ya));
    const double yc = __ddiv_rn(__dsub_rn(__dsub_rn(__dmul_rn(2.0, y), ya), yb), __dsub_rn(yb, ya));
    const double y2 = __dmul_rn(2.0, yc);
    for (int64_t k = 0; k < N; ++k) col[k * S2_LD] = 0.0;
    // four functions of f.A at a time: four independent Clenshaw chains (ILP), then ONE shared-memory read-modify-write
    // per coefficient k that applies the four rank-one updates in r order (the same fma sequence as the scalar loop)
    constexpr int RB = 4;
    for (int64_t r0 = 0; r0 < R; r0 += RB) {
        double fa[RB];
        double b1[RB], b2[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i) { fa[i] = 0.0; b1[i] = 0.0; b2[i] = 0.0; }
        if (inside && NA >= 1) {
            for (int64_t k = NA - 1; k >= 1; --k) {
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const double ak = (r0 + i < R) ? __ldg(A + (r0 + i) * NA + k) : 0.0;
                    const double t = __fma_rn(y2, b1[i], __dsub_rn(ak, b2[i]));
                    b2[i] = b1[i];
                    b1[i] = t;
                }
            }
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const double a0 = (r0 + i < R) ? __ldg(A + (r0 + i) * NA) : 0.0;
                fa[i] = (NA == 1) ? a0 : __fma_rn(__dmul_rn(y2, 0.5), b1[i], __dsub_rn(a0, b2[i]));
            }
        }
        for (int64_t k = 0; k < N; ++k) {
            double acc = col[k * S2_LD];
#pragma unroll
            for (int i = 0; i < RB; ++i)
                if (r0 + i < R) acc = __fma_rn(__ldg(CB + (r0 + i) * N + k), fa[i], acc);
            col[k * S2_LD] = acc;
        }
    }
    // chebnormalizedcumsum! (matrix semantics: no growth), same operation order as normcumsum_kernel
    const int64_t n = N;
    if (n == 2) {
        col[S2_LD] = __dmul_rn(col[S2_LD], 0.5);
    } else if (n > 2) {
        col[0] = __dsub_rn(col[0], __dmul_rn(col[2 * S2_LD], 0.5));
        for (int64_t q = 1; q <= n - 3; ++q) col[q * S2_LD] = __dmul_rn(__dsub_rn(col[q * S2_LD], col[(q + 2) * S2_LD]), 0.5);
        col[(n - 2) * S2_LD] = __dmul_rn(col[(n - 2) * S2_LD], 0.5);
        col[(n - 1) * S2_LD] = __dmul_rn(col[(n - 1) * S2_LD], 0.5);
    }
    for (int64_t k = n - 1; k >= 1; --k) col[k * S2_LD] = __ddiv_rn(col[(k - 1) * S2_LD], (double)k);
    if (n > 0) col[0] = 0.0;
    for (int64_t k = 2; k <= n; ++k) col[0] = __dadd_rn(col[0], (k & 1) ? -col[(k - 1) * S2_LD] : col[(k - 1) * S2_LD]);
    double val = 0.0;
    for (int64_t k = 0; k < n; ++k) val = __dadd_rn(val, col[k * S2_LD]);
    val = __ddiv_rn(1.0, val);
    for (int64_t k = 0; k < n; ++k) col[k * S2_LD] = __dmul_rn(col[k * S2_LD], val);
    // 47 Clenshaw bisections on the thread's own column
    double lo = -1.0, hi = 1.0;
    const double uu = act ? u[j] : 0.5;
    for (int it = 0; it < numits; ++it) {
        const double m = __dmul_rn(0.5, __dadd_rn(lo, hi));
        double v;
        if (n == 0) v = 0.0;
        else if (n == 1) v = col[0];
        else {
            const double x2 = __dmul_rn(2.0, m);
            double b1 = 0.0, b2 = 0.0;
            for (int64_t k = n - 1; k >= 1; --k) {
                const double t = __fma_rn(x2, b1, __dsub_rn(col[k * S2_LD], b2));
                b2 = b1;
                b1 = t;
            }
            v = __fma_rn(__dmul_rn(x2, 0.5), b1, __dsub_rn(col[0], b2));
        }
        if (v <= uu) lo = m; else hi = m;
    }
    const double xm = __dmul_rn(0.5, __dadd_rn(lo, hi));
    const double mid = __dmul_rn(__dadd_rn(xa, xb), 0.5), len = __dsub_rn(xb, xa);
    if (act) rx[j] = __dadd_rn(mid, __dmul_rn(__dmul_rn(len, xm), 0.5));
}

// unfused fallback piece for very long B factors (tile does not fit shared memory): AB = CB * fA
__global__ void gemm_nn_kernel(const double* __restrict__ CB, const double* __restrict__ fA, double* __restrict__ AB,
                               int64_t N, int64_t R, int64_t P) {
    const int64_t total = N * P;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = g % N, j = g / N;
        double acc = 0.0;
        for (int64_t r = 0; r < R; ++r) acc = __fma_rn(CB[k + N * r], fA[r + R * j], acc);
        AB[g] = acc;
    }
}

// ---- unfused route of the 2-D conditional sampler (any N): the reference's own sequence of src/Extras/sample.jl:208-213 as
// separate kernels over chunks of samples -- evaluate.(f.A, ry') -> CB*fA -> chebnormalizedcumsum! -> chebbisectioninv ->
// fromcanonical -- with the operation order of the fused kernel, so both routes give the same bits.
__global__ void s2_tocanonical_kernel(const double* __restrict__ ry, double* __restrict__ yc, int64_t P, double ya, double yb) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x)
        yc[j] = __ddiv_rn(__dsub_rn(__dsub_rn(__dmul_rn(2.0, ry[j]), ya), yb), __dsub_rn(yb, ya));
}
// evaluate returns zero outside the domain
__global__ void s2_mask_outside_kernel(const double* __restrict__ ry, double* __restrict__ fA, int64_t R, int64_t P, double lo, double hi) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < R * P; g += (int64_t)gridDim.x * blockDim.x) {
        const double y = ry[g / R];
        if (!(y >= lo && y <= hi)) fA[g] = 0.0;
    }
}
__global__ void s2_fromcanonical_kernel(double* __restrict__ rx, int64_t P, double xa, double xb) {
    const double mid = __dmul_rn(__dadd_rn(xa, xb), 0.5), len = __dsub_rn(xb, xa);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x)
        rx[j] = __dadd_rn(mid, __dmul_rn(__dmul_rn(len, rx[j]), 0.5));
}
static int32_t sample2d_unfused(const double* d_A, int64_t NA, const double* d_CB, int64_t N, int64_t R, double ya, double yb,
                                double xa, double xb, const double* d_ry, const double* d_u, double* d_rx, int64_t P, int numits,
                                void* stream) {
    const int64_t per = 8 * (N + R + 1);
    int64_t chunk = std::max<int64_t>(1024, ((int64_t)256 << 20) / per);
    if (chunk > P) chunk = P;
    double *yc = nullptr, *fA = nullptr, *AB = nullptr;
    void* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)(chunk * (N + R + 1))));
    yc = reinterpret_cast<double*>(d); fA = yc + chunk; AB = fA + chunk * R;
    int32_t st = AFB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t p0 = 0; p0 < P && st == AFB_OK; p0 += chunk) {
        const int64_t np = std::min(chunk, P - p0);
        int64_t grid = std::min<int64_t>((np + 255) / 256, 8 * (int64_t)sm_count());
        auto k1 = s2_tocanonical_kernel;
        AFB_LAUNCH(k1, (unsigned)grid, 256, 0, s, d_ry + p0, yc, np, ya, yb);
        st = launch_multi(d_A, NA, R, yc, fA, np, stream);
        if (st != AFB_OK) break;
        auto k2 = s2_mask_outside_kernel;
        grid = std::min<int64_t>((np * R + 255) / 256, 8 * (int64_t)sm_count());
        if (R > 0) AFB_LAUNCH(k2, (unsigned)grid, 256, 0, s, d_ry + p0, fA, R, np, ya < yb ? ya : yb, ya < yb ? yb : ya);
        if (R > 0) st = launch_gemm_nn(d_CB, fA, AB, N, R, np, stream);
        else if (N * np > 0) { cudaError_t e = cudaMemsetAsync(AB, 0, sizeof(double) * (size_t)(N * np), s); if (e != cudaSuccess) st = cuda_fail(e, "memset", __FILE__, __LINE__); }
        if (st == AFB_OK) st = launch_normcumsum(AB, N, np, N, 0, stream);
        if (st == AFB_OK) st = launch_cols(true, AB, N, N, d_u + p0, d_rx + p0, np, numits, stream);
        if (st != AFB_OK) break;
        auto k3 = s2_fromcanonical_kernel;
        grid = std::min<int64_t>((np + 255) / 256, 8 * (int64_t)sm_count());
        AFB_LAUNCH(k3, (unsigned)grid, 256, 0, s, d_rx + p0, np, xa, xb);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && st == AFB_OK) st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    cudaFree(d);
    if (st == AFB_OK) { e = cudaGetLastError(); if (e != cudaSuccess) st = cuda_fail(e, "sample2d_unfused", __FILE__, __LINE__); }
    return st;
}

int32_t launch_sample2d(const double* d_A, int64_t NA, const double* d_CB, int64_t N, int64_t R, double ya, double yb,
                        double xa, double xb, const double* d_ry, const double* d_u, double* d_rx, int64_t P, int numits,
                        void* stream) {
    if (P == 0) return AFB_OK;
    const size_t smem = sizeof(double) * (size_t)(N > 0 ? N : 1) * S2_LD;
    if (smem > 200 * 1024)   // B factors longer than 198 coefficients: the column tile no longer fits shared memory
        return sample2d_unfused(d_A, NA, d_CB, N, R, ya, yb, xa, xb, d_ry, d_u, d_rx, P, numits, stream);
    auto k = sample2d_kernel;
    AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AFB_LAUNCH(k, (unsigned)((P + S2_W - 1) / S2_W), S2_W, smem, stream, d_A, NA, d_CB, N, R, ya, yb, xa, xb, d_ry, d_u, d_rx,
               P, numits);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// FP64 tensor-core instruction (DMMA): D(8x8) = A(8x4, row) * B(4x8, col) + C.  Per thread: a = A[lane/4][lane%4],
// b = B[lane%4][lane/4], c/d = {C[lane/4][2(lane%4)], C[lane/4][2(lane%4)+1]}.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
#ifndef AFB_HOST_EMU
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
#else
    (void)a; (void)b; (void)d0; (void)d1;
#endif
}
// The same contraction on the FP64 tensor cores (DMMA m8n8k4), written to MEASURE the tensor-core route BASELINE.json asks
// about (DESIGN.md 4b has the numbers): a warp owns 32 samples (four 8-sample A fragments of fA^T, loaded once) and walks the
// N coefficients in tiles of 8; CB sits in shared memory (one LDS.64 feeds four DMMAs); D = AB^T tiles are stored as 16-byte
// pairs of consecutive coefficients.  The accumulation order inside a DMMA is the hardware's, not the scalar loop's, so the
// result agrees with gemm_nn_kernel to rounding only -- the product path of the sampler keeps the bit-exact FMA kernels.
constexpr int DG_WARPS = 8, DG_MI = 4, DG_KT = 8;
__global__ void __launch_bounds__(32 * DG_WARPS) gemm_nn_dmma_kernel(const double* __restrict__ CB, const double* __restrict__ fA,
                                                                     double* __restrict__ AB, int64_t N, int64_t R, int64_t P) {
    AFB_DYN_SMEM(double, sCB);   // Np x Rp, zero padded to multiples of 8 and 4, row k contiguous in r (+1 pad)
    const int64_t Np = (N + 7) & ~(int64_t)7, Rp = (R + 3) & ~(int64_t)3;
    const int ldc = (int)Rp + 1;
    for (int64_t g = threadIdx.x; g < Np * Rp; g += blockDim.x) {
        const int64_t k = g / Rp, r = g % Rp;
        sCB[k * ldc + r] = (k < N && r < R) ? CB[k + N * r] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int64_t nkt = Rp / 4;
    for (int64_t j0 = ((int64_t)blockIdx.x * DG_WARPS + warp) * (8 * DG_MI); j0 < P; j0 += (int64_t)gridDim.x * DG_WARPS * 8 * DG_MI) {
        if (nkt <= DG_KT) {
            // the warp's fA fragments stay in registers while it walks the coefficient tiles (R <= 32)
            double a[DG_MI][DG_KT];
#pragma unroll
            for (int mi = 0; mi < DG_MI; ++mi)
#pragma unroll
                for (int kt = 0; kt < DG_KT; ++kt) {
                    const int64_t j = j0 + 8 * mi + grp, r = 4 * kt + tig;
                    a[mi][kt] = (kt < nkt && j < P && r < R) ? __ldg(fA + r + R * j) : 0.0;
                }
            for (int64_t nt = 0; nt < Np / 8; ++nt) {
                double d[DG_MI][2];
#pragma unroll
                for (int mi = 0; mi < DG_MI; ++mi) { d[mi][0] = 0.0; d[mi][1] = 0.0; }
                const double* brow = sCB + (8 * nt + grp) * ldc + tig;
#pragma unroll
                for (int kt = 0; kt < DG_KT; ++kt) {
                    if (kt < nkt) {
                        const double b = brow[4 * kt];
#pragma unroll
                        for (int mi = 0; mi < DG_MI; ++mi) dmma_m8n8k4(d[mi][0], d[mi][1], a[mi][kt], b);
                    }
                }
#pragma unroll
                for (int mi = 0; mi < DG_MI; ++mi) {
                    const int64_t j = j0 + 8 * mi + grp, k = 8 * nt + 2 * tig;
                    if (j < P) {
                        if (k + 1 < N && (N & 1) == 0) *reinterpret_cast<double2*>(AB + k + N * j) = make_double2(d[mi][0], d[mi][1]);
                        else { if (k < N) AB[k + N * j] = d[mi][0]; if (k + 1 < N) AB[k + 1 + N * j] = d[mi][1]; }
                    }
                }
            }
            continue;
        }
        for (int64_t nt = 0; nt < Np / 8; ++nt) {
            double d[DG_MI][2];
#pragma unroll
            for (int mi = 0; mi < DG_MI; ++mi) { d[mi][0] = 0.0; d[mi][1] = 0.0; }
            for (int64_t kt = 0; kt < nkt; ++kt) {
                const int64_t r = 4 * kt + tig;
                const double b = sCB[(8 * nt + grp) * ldc + r];            // B[k = r][n = coefficient 8 nt + grp]
#pragma unroll
                for (int mi = 0; mi < DG_MI; ++mi) {
                    const int64_t j = j0 + 8 * mi + grp;                    // A[m = sample][k = r] = fA[r, j]
                    const double a = (j < P && r < R) ? __ldg(fA + r + R * j) : 0.0;
                    dmma_m8n8k4(d[mi][0], d[mi][1], a, b);
                }
            }
#pragma unroll
            for (int mi = 0; mi < DG_MI; ++mi) {
                const int64_t j = j0 + 8 * mi + grp, k = 8 * nt + 2 * tig;
                if (j < P) {
                    if (k < N) AB[k + N * j] = d[mi][0];
                    if (k + 1 < N) AB[k + 1 + N * j] = d[mi][1];
                }
            }
        }
    }
}
int32_t launch_gemm_nn_dmma(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P, void* stream) {
    if (N * P == 0) return AFB_OK;
    const int64_t Np = (N + 7) & ~(int64_t)7, Rp = (R + 3) & ~(int64_t)3;
    const size_t smem = sizeof(double) * (size_t)(Np * (Rp + 1));
    if (smem > 200 * 1024) return fail(AFB_UNSUPPORTED, "DMMA GEMM: CB does not fit shared memory");
    int64_t grid = (P + DG_WARPS * 8 * DG_MI - 1) / (DG_WARPS * 8 * DG_MI);
    if (grid > 8 * (int64_t)sm_count()) grid = 8 * (int64_t)sm_count();
    auto k = gemm_nn_dmma_kernel;
    AFB_SMEM_OPTIN(k, 200 * 1024);
    AFB_LAUNCH(k, (unsigned)grid, 32 * DG_WARPS, smem, stream, d_CB, d_fA, d_AB, N, R, P);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_gemm_nn(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P, void* stream) {
    if (N * P == 0) return AFB_OK;
    int64_t grid = (N * P + 255) / 256;
    if (grid > 16 * (int64_t)sm_count()) grid = 16 * (int64_t)sm_count();
    auto k = gemm_nn_kernel;
    AFB_LAUNCH(k, (unsigned)grid, 256, 0, stream, d_CB, d_fA, d_AB, N, R, P);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// ------------------------------------------------------------------------------------------------
// FP64 issue-rate probe (measurement only, used by bench.py to replace the nominal FP64 peak by a measured one and to
// test the register-file-port explanation of the Clenshaw ceiling).  Every thread runs 8 independent
// dependency chains for `iters` steps.
//   mode 0: DFMA with three distinct register operands          x = fma(a, x, c)        (a, c per-chain registers)
//   mode 1: DFMA with a repeated register operand                x = fma(x, x, c)
//   mode 2: the Clenshaw step  t = c - b2 ; b = fma(x2, b1, t)   (DADD + 3-operand DFMA)
//   mode 3: reassociated step  s = fma(x2, b1, c_uniform) ; b = s - b2  (2-register DFMA + DADD; different rounding)
// flops counted by the caller: modes 0/1: 2 per step and chain, mode 2: 3.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(int mode, int iters, double seed, double* __restrict__ out) {
    if (mode == 4) {
        // DMMA issue rate: 8 independent accumulator tiles per warp, 512 flop per instruction and warp
        double d[8][2];
        double av = 0.999 + 1e-9 * (double)threadIdx.x, bv = 1e-3 + 1e-12 * (double)threadIdx.x;
#pragma unroll
        for (int i = 0; i < 8; ++i) { d[i][0] = seed * (double)(i + 1); d[i][1] = seed * (double)(i + 9); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma_m8n8k4(d[i][0], d[i][1], av, bv);
        }
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += d[i][0] + d[i][1];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
        return;
    }
    double x[8], a[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = seed * (double)(threadIdx.x + 1 + i);
        a[i] = 0.999999 + 1e-9 * (double)(i + threadIdx.x);
        c[i] = 1e-3 * (double)(i + 1) + 1e-12 * (double)threadIdx.x;   // per-thread: keeps it out of the uniform registers
    }
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fma_rn(a[i], x[i], c[i]);
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fma_rn(x[i], x[i], c[i]);
        }
    } else if (mode == 2) {
        double b2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) b2[i] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double t = __fma_rn(a[i], x[i], __dsub_rn(c[i], b2[i]));
                b2[i] = x[i];
                x[i] = t;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += b2[i];
    } else {
        // mode 3: reassociated step with a warp-uniform coefficient  s = fma(x2, b1, c_uniform) ; b = s - b2
        double b2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) b2[i] = 0.0;
        for (int it = 0; it < iters; ++it) {
            const double cu = seed * (double)(it & 1023);   // uniform across the warp: lives in a uniform register
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double t = __dsub_rn(__fma_rn(a[i], x[i], cu), b2[i]);
                b2[i] = x[i];
                x[i] = t;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += b2[i];
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int32_t launch_fp64_probe(int mode, int iters, int blocks, double* d_out, void* stream) {
    auto k = fp64_probe_kernel;
    AFB_LAUNCH(k, (unsigned)blocks, 256, 0, stream, mode, iters, 1e-7, d_out);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

}  // namespace afb
