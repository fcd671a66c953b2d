// Any-length batched complex FFT engine and the generic (non fused) real-transform path.
//
//   * power-of-two n <= 8192      : one shared-memory line kernel (afb_transform.cu)
//   * power-of-two n  > 8192      : four-step  n = n1 n2  (kernel K3 of SURVEY.md section 2c):
//        A: n2 strided lines of length n1 (+ twiddle w_n^(b k1)),  B: n1 contiguous lines of length n2
//           with a stride-n1 scatter so that X[k1 + n1 k2] lands in natural order
//   * any other n                 : Bluestein chirp-z through a power-of-two engine of length M >= 2n-1
//
// The generic real-transform path (plot-sized or chopped lengths, and n > 16384) maps each kind onto
// a length-n complex FFT with an element-wise pre and post kernel; it is the correctness path for
// arbitrary n -- the fused power-of-two kernels of afb_transform.cu are the throughput path.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "afb_fft.cuh"
#include "afb_internal.h"

namespace afb {

struct CfftPlan {
    int64_t n = 0, max_batch = 0;
    int mode = 0;  // 0 line, 1 four-step, 2 Bluestein, 3 mixed radix (2^a 3^b 5^c 7^d in one CTA, afb_mixed.cu),
                   // 4 split: one radix-S decimation-in-frequency pass (S = 2, 3, 4, 5, 7, 8, 16) in front of S one-CTA transforms of n / S
    const cplx* tw = nullptr;
    // four-step
    int n1 = 0, n2 = 0;
    const cplx *tw1 = nullptr, *tw2 = nullptr;
    cplx *twA = nullptr, *twB = nullptr;
    cplx* tmp = nullptr;
    // split
    int S = 1;
    const cplx* rootn = nullptr;   // exp(-2 pi i k / n), k < n
    // Bluestein (sub: the padded power-of-two engine) / split (sub: the engine of length n / S)
    int64_t M = 0;
    CfftPlan* sub = nullptr;
    cplx *chirp = nullptr, *bhat = nullptr, *pad = nullptr;
    int bS = 1;                     // Bluestein with M = bS * 8192 > 8192: bhat holds bS blocks (element bS k + s at s h + k)
    const cplx* rootM = nullptr;    // exp(-2 pi i m / M)
};

// ---- Bluestein element-wise kernels ------------------------------------------------------------------
__global__ void blue_pre_kernel(const cplx* __restrict__ in, cplx* __restrict__ pad, int64_t n, int64_t M,
                                int64_t batch, const cplx* __restrict__ chirp, int conj_in) {
    const int64_t total = batch * M;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / M, j = g % M;
        cplx z = make_double2(0.0, 0.0);
        if (j < n) {
            z = in[b * n + j];
            if (conj_in) z.y = -z.y;
            z = cmul(z, chirp[j]);
        }
        pad[g] = z;
    }
}
__global__ void blue_mul_kernel(cplx* __restrict__ pad, int64_t M, int64_t batch, const cplx* __restrict__ bhat) {
    const int64_t total = batch * M;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x)
        pad[g] = cmul(pad[g], bhat[g % M]);
}
__global__ void blue_post_kernel(const cplx* __restrict__ pad, cplx* __restrict__ out, int64_t n, int64_t M,
                                 int64_t batch, const cplx* __restrict__ chirp, int conj_out, double scale) {
    const int64_t total = batch * n;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / n, k = g % n;
        cplx z = cscale(cmul(pad[b * M + k], chirp[k]), scale);
        if (conj_out) z.y = -z.y;
        out[g] = z;
    }
}
__global__ void fill_bfilter_kernel(cplx* __restrict__ b, int64_t n, int64_t M, const cplx* __restrict__ chirp) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        cplx z = make_double2(0.0, 0.0);
        if (j < n) z = cconj(chirp[j]);
        else if (M - j < n) z = cconj(chirp[M - j]);
        b[j] = z;
    }
}

// ---- Bluestein in ONE kernel when the padded length M fits a CTA's shared memory (M <= 8192) ---------------------
// chirp-multiply and zero-pad in registers -> FFT_M in shared memory -> multiply by bhat -> inverse FFT_M (conj trick)
// -> chirp-multiply -> store: one read and one write of the line instead of five element-wise / FFT passes over a
// padded copy.
struct BlueParams {
    const cplx* in;
    cplx* out;
    int64_t n, batch;
    const cplx *chirp, *bhat, *tw;   // chirp == nullptr: plain circular convolution core (FFT, filter, inverse FFT)
    int conj_in, conj_out;
    double scale;  // includes 1/M
    int S = 1;     // filters: line L uses bhat + (L % S) M
};
template <int M> struct BlueGeom {
    static constexpr int T = FftCfg<M>::T;
    static constexpr int CT = (T < 128) ? 128 : T;
    static constexpr int LPC = CT / T;
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)LPC * FftLine<M>::STRIDE;
};
template <int M>
__global__ void __launch_bounds__(BlueGeom<M>::CT, M == 8192 ? 1 : (M == 4096 && AFB_R32_4096 ? 2 : 512 / BlueGeom<M>::CT)) blue_fused_kernel(const __grid_constant__ BlueParams p) {
    typedef FftCfg<M> C;
    typedef BlueGeom<M> G;
    AFB_DYN_SMEM(cplx, smem);
    const int l = threadIdx.x / C::T, t = threadIdx.x % C::T;
    const int64_t line = (int64_t)blockIdx.x * G::LPC + l;
    const bool act = line < p.batch;
    cplx* s = smem + (size_t)l * FftLine<M>::STRIDE;
    const cplx* src = p.in + (act ? line : 0) * p.n;
    cplx x[C::E];
#pragma unroll
    for (int r = 0; r < C::E; ++r) {  // unconditional loads on a clamped index, then select
        const int64_t j = t + C::T * r;
        const int64_t jc = j < p.n ? j : p.n - 1;
        cplx z = src[jc];
        if (p.conj_in) z.y = -z.y;
        if (p.chirp) z = cmul(z, ldg2(p.chirp + jc));
        x[r] = (j < p.n) ? z : make_double2(0.0, 0.0);
    }
    fft_line<M>(s, t, x, p.tw);
    const cplx* bh = p.bhat + (p.S > 1 ? (line % p.S) * M : 0);
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int i = t + C::T * r;
        const cplx z = cmul(line_get<M>(s, i), ldg2(bh + i));
        x[r] = make_double2(z.x, -z.y);  // inverse FFT = conj FFT conj
    }
    __syncthreads();  // every read of the first result is done before the second transform reuses the buffer
    // The second transform has the same thread index, buffer and twiddles: hide that from the compiler, which otherwise
    // keeps the first transform's shared-memory addresses and twiddle powers alive for reuse (~100 registers, spilled).
    int t2 = t;
    const cplx* tw2 = p.tw;
    cplx* s2 = s;
#ifndef AFB_HOST_EMU
    asm volatile("" : "+r"(t2), "+l"(tw2), "+l"(s2));
#endif
    fft_line<M>(s2, t2, x, tw2);
    if (!act) return;
    cplx* dst = p.out + line * p.n;
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int64_t k = t2 + C::T * r;
        if (k < p.n) {
            const cplx y = line_get<M>(s2, (int)k);
            cplx z = make_double2(y.x, -y.y);
            if (p.chirp) z = cmul(z, ldg2(p.chirp + k));
            z = cscale(z, p.scale);
            if (p.conj_out) z.y = -z.y;
            dst[k] = z;
        }
    }
}
template <int M> static int32_t launch_blue_fused_m(const BlueParams& p, void* stream) {
    typedef BlueGeom<M> G;
    auto k = blue_fused_kernel<M>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    AFB_LAUNCH(k, (unsigned)((p.batch + G::LPC - 1) / G::LPC), G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
static int32_t launch_blue_fused(int64_t M, const BlueParams& p, void* stream) {
    switch (M) {
        case 32: return launch_blue_fused_m<32>(p, stream);
        case 64: return launch_blue_fused_m<64>(p, stream);
        case 128: return launch_blue_fused_m<128>(p, stream);
        case 256: return launch_blue_fused_m<256>(p, stream);
        case 512: return launch_blue_fused_m<512>(p, stream);
        case 1024: return launch_blue_fused_m<1024>(p, stream);
        case 2048: return launch_blue_fused_m<2048>(p, stream);
        case 4096: return launch_blue_fused_m<4096>(p, stream);
        case 8192: return launch_blue_fused_m<8192>(p, stream);
        default: return fail(AFB_UNSUPPORTED, "fused Bluestein: padded length not instantiated");
    }
}

static inline unsigned ew_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    const int64_t cap = 16 * (int64_t)sm_count();
    return (unsigned)std::max<int64_t>(1, std::min(g, cap));
}

template <int S>
__global__ void blue_split_pre_kernel(const cplx* __restrict__ in, cplx* __restrict__ pad, int64_t n, int64_t h, int64_t batch,
                                      const cplx* __restrict__ chirp, const cplx* __restrict__ rootM, int conj_in);
template <int S>
__global__ void blue_split_post_kernel(const cplx* __restrict__ pad, cplx* __restrict__ out, int64_t n, int64_t h, int64_t batch,
                                       const cplx* __restrict__ chirp, const cplx* __restrict__ rootM, int conj_out, double scale);

// lengths the one-CTA engines do not reach (smooth n > 6912, Bluestein with a padded length above 8192) but whose S-th part
// they do: the smallest S in {2, 4, 8, 16} with n / S served by the mixed-radix kernel or the fused Bluestein kernel
static bool one_cta_engine_ok(int64_t h) {
    if (is_pow2(h)) return h >= LINE_MIN_N && h <= LINE_MAX_N;
    return mixed_radix_ok(h) || (h >= 2 && 2 * h - 1 <= LINE_MAX_N);
}
// Cost model behind the choice of S (fitted to one box, 4.3 GB batches, ms: n = 8000 S = 2 / 4 / 8 -> 4.87 / 4.41 / 4.66;
// 13 824 S = 4 / 8 / 16 -> 4.24 / 4.95 / 7.41; 16 000 S = 4 / 8 / 16 -> 5.06 / 5.31 / 6.98; 40 000 S = 8 / 16 -> 5.92 / 7.43):
// the radix-S pass costs O(S) complex multiplies per element (S = 16 is expensive), the one-CTA transform gets cheaper per
// element as h shrinks (more resident CTAs), a Bluestein block costs in proportion to its padding M / h.
static double split_cost(int S, int64_t h) {
    static const double pre[17] = {0, 0, 0.0, 0.05, 0.1, 0.3, 0, 0.6, 0.9, 0, 0, 0, 0, 0, 0, 0, 2.7};
    double mid;
    if (is_pow2(h)) mid = 0.7;
    else if (mixed_radix_ok(h)) mid = h <= 1024 ? 1.0 : h <= 2048 ? 1.5 : h <= 4096 ? 2.05 : 3.0;
    else {
        int64_t M = 32;
        while (M < 2 * h - 1) M <<= 1;
        mid = 2.35 * (double)M / (double)h;
    }
    return pre[S] + mid;
}
static int split_factor(int64_t n) {
    if (n < 2 || is_pow2(n) || one_cta_engine_ok(n)) return 1;
    static const int kS[] = {2, 3, 4, 5, 7, 8, 16};   // odd S serve smooth odd lengths: 7875 = 3 x 2625, 16807 = 7 x 2401
    int best = 1;
    double cbest = 0.0;
    for (int S : kS) {
        if (n % S != 0 || !one_cta_engine_ok(n / S)) continue;
        const double c = split_cost(S, n / S);
        if (best == 1 || c < cbest) { best = S; cbest = c; }
    }
    return best;
}
int cfft_split(const CfftPlan* p) { return (p && p->mode == 4) ? p->S : 1; }

void cfft_destroy(CfftPlan* p) {
    if (!p) return;
    if (p->twA) cudaFree(p->twA);
    if (p->twB) cudaFree(p->twB);
    if (p->tmp) cudaFree(p->tmp);
    if (p->chirp) cudaFree(p->chirp);
    if (p->bhat) cudaFree(p->bhat);
    if (p->pad) cudaFree(p->pad);
    cfft_destroy(p->sub);
    delete p;
}

int32_t cfft_create(CfftPlan** out, int64_t n, int64_t max_batch) {
    CfftPlan* p = new CfftPlan();
    p->n = n;
    p->max_batch = std::max<int64_t>(1, max_batch);
    int32_t st = AFB_OK;
    const long double PI = 3.141592653589793238462643383279502884L;
    if (is_pow2(n) && n >= LINE_MIN_N && n <= LINE_MAX_N) {
        p->mode = 0;
        st = root_table(n, &p->tw);
    } else if (is_pow2(n) && n > LINE_MAX_N) {
        const int lg = ilog2(n);
        if (lg > 26) { cfft_destroy(p); return fail(AFB_UNSUPPORTED, "complex FFT length above 2^26"); }
        p->mode = 1;
        p->n1 = 1 << ((lg + 1) / 2);
        p->n2 = 1 << (lg / 2);
        st = root_table(p->n1, &p->tw1);
        if (st == AFB_OK) st = root_table(p->n2, &p->tw2);
        if (st == AFB_OK) {
            std::vector<cplx> hA((size_t)((n + 4095) / 4096)), hB(4096);
            for (size_t i = 0; i < hA.size(); ++i) {
                const long double a = 2.0L * PI * (long double)(i * 4096) / (long double)n;
                hA[i] = make_double2((double)cosl(a), (double)(-sinl(a)));
            }
            for (size_t i = 0; i < 4096; ++i) {
                const long double a = 2.0L * PI * (long double)i / (long double)n;
                hB[i] = make_double2((double)cosl(a), (double)(-sinl(a)));
            }
            st = upload_table(hA, &p->twA);
            if (st == AFB_OK) st = upload_table(hB, &p->twB);
        }
        if (st == AFB_OK) {
            void* d = nullptr;
            cudaError_t e = cudaMalloc(&d, sizeof(cplx) * (size_t)(n * p->max_batch));
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(four-step tmp)", __FILE__, __LINE__);
            p->tmp = reinterpret_cast<cplx*>(d);
        }
    } else if (mixed_radix_ok(n)) {
        p->mode = 3;
        st = root_table(n, &p->tw);
    } else if (split_factor(n) > 1) {
        p->mode = 4;
        p->S = split_factor(n);
        st = root_table(n, &p->rootn);
        if (st == AFB_OK) st = cfft_create(&p->sub, n / p->S, p->max_batch);
    } else if (n >= 1) {
        p->mode = 2;
        int64_t M = 32;
        while (M < 2 * n - 1) M <<= 1;
        p->M = M;
        // M above one CTA and at most 16 x 8192: FFT_M as a radix-bS pass around length-8192 one-CTA transforms (see
        // blue_split_pre_kernel); the run-time engine is then the 8192 line plan, a length-M plan only builds the filter
        const bool bsplit = M > LINE_MAX_N && M / LINE_MAX_N <= 16;
        CfftPlan* full = nullptr;   // length-M engine
        st = cfft_create(&full, M, bsplit ? 1 : p->max_batch);
        if (st == AFB_OK) {
            std::vector<cplx> h((size_t)n);
            for (int64_t j = 0; j < n; ++j) {
                const int64_t m = (int64_t)(((__int128)j * j) % (2 * n));  // j^2 mod 2n, exact
                const long double a = PI * (long double)m / (long double)n;
                h[(size_t)j] = make_double2((double)cosl(a), (double)(-sinl(a)));  // exp(-i pi j^2 / n)
            }
            st = upload_table(h, &p->chirp);
        }
        if (st == AFB_OK) {
            void* d = nullptr;
            cudaError_t e = cudaMalloc(&d, sizeof(cplx) * (size_t)M);
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(bhat)", __FILE__, __LINE__);
            p->bhat = reinterpret_cast<cplx*>(d);
        }
        if (st == AFB_OK) {
            void* d = nullptr;
            // the padded work buffer is only needed by the multi-kernel route (M above one CTA's shared memory)
            cudaError_t e = (M <= LINE_MAX_N) ? cudaSuccess : cudaMalloc(&d, sizeof(cplx) * (size_t)(M * p->max_batch));
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMalloc(bluestein pad)", __FILE__, __LINE__);
            p->pad = reinterpret_cast<cplx*>(d);
        }
        if (st == AFB_OK) {
            auto k = fill_bfilter_kernel;
            AFB_LAUNCH(k, ew_grid(M), 256, 0, nullptr, p->bhat, n, M, p->chirp);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) st = cuda_fail(e, "fill_bfilter_kernel", __FILE__, __LINE__);
            if (st == AFB_OK) st = cfft_exec(full, p->bhat, p->bhat, 1, false, 1.0, nullptr);
            if (st == AFB_OK) {
                e = cudaStreamSynchronize(nullptr);
                if (e != cudaSuccess) st = cuda_fail(e, "sync(bhat)", __FILE__, __LINE__);
            }
        }
        if (st == AFB_OK && bsplit) {
            // filter blocks: element bS k + s of the spectrum moves to s h + k
            const int S = (int)(M / LINE_MAX_N);
            const int64_t h = LINE_MAX_N;
            std::vector<cplx> nat((size_t)M), blk((size_t)M);
            cudaError_t e = cudaMemcpy(nat.data(), p->bhat, sizeof(cplx) * (size_t)M, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) st = cuda_fail(e, "cudaMemcpy(bhat)", __FILE__, __LINE__);
            for (int64_t k = 0; k < h; ++k)
                for (int sgm = 0; sgm < S; ++sgm) blk[(size_t)(sgm * h + k)] = nat[(size_t)(S * k + sgm)];
            if (st == AFB_OK) {
                e = cudaMemcpy(p->bhat, blk.data(), sizeof(cplx) * (size_t)M, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) st = cuda_fail(e, "cudaMemcpy(bhat blocks)", __FILE__, __LINE__);
            }
            if (st == AFB_OK) st = root_table(M, &p->rootM);
            if (st == AFB_OK) st = cfft_create(&p->sub, h, 1);
            p->bS = S;
            cfft_destroy(full);
        } else {
            p->sub = full;
        }
    } else {
        st = fail(AFB_BAD_SIZE, "complex FFT length must be >= 1");
    }
    if (st != AFB_OK) { cfft_destroy(p); return st; }
    *out = p;
    return AFB_OK;
}

// the transform(s) of a work buffer filled by launch_generic_pre / launch_ext_pre (S blocks per line for a split plan)
int32_t cfft_exec_work(CfftPlan* p, cplx* work, int64_t nfft, bool inverse, void* stream) {
    if (p->mode == 4) return cfft_exec(p->sub, work, work, nfft * p->S, inverse, 1.0, stream);
    return cfft_exec(p, work, work, nfft, inverse, 1.0, stream);
}
int32_t cfft_work_launches(const CfftPlan* p) { return (p && p->mode == 4) ? cfft_launches(p->sub) : cfft_launches(p); }

int32_t cfft_launches(const CfftPlan* p) {
    if (!p) return 0;
    switch (p->mode) {
        case 0: return 1;
        case 1: return 2;
        case 3: return 1;
        case 4: return cfft_launches(p->sub);
        default: return (p->M <= LINE_MAX_N) ? 1 : (p->bS > 1 ? 3 : 3 + 2 * cfft_launches(p->sub));
    }
}

int32_t cfft_exec(CfftPlan* p, const cplx* in, cplx* out, int64_t batch, bool inverse, double scale, void* stream) {
    if (batch == 0) return AFB_OK;
    const int64_t n = p->n;
    if (p->mode == 0) {
        return launch_cfft_lines((int)n, in, out, batch, n, 1, n, 1, p->tw, inverse, inverse, scale, nullptr, nullptr, 1,
                                 stream);
    }
    if (p->mode == 1) {
        const int64_t n1 = p->n1, n2 = p->n2;
        // whole chunks of transforms per launch: line L = (b, column or row) through the lpb / ibs / obs addressing
        for (int64_t b0 = 0; b0 < batch; b0 += p->max_batch) {
            const int64_t nb = std::min(p->max_batch, batch - b0);
            const cplx* src = in + b0 * n;
            cplx* dst = out + b0 * n;
            // A: for every column b2 (stride-n2 elements): FFT over a, times w_n^(b2 k1), kept in place layout
            AFB_TRY(launch_cfft_lines((int)n1, src, p->tmp, nb * n2, 1, n2, 1, n2, p->tw1, inverse, false, 1.0, p->twA,
                                      p->twB, n2, stream, 0, 0, n2, n, n));
            // B: for every row k1 (contiguous n2 elements): FFT over b2, scattered to X[k1 + n1 k2]
            AFB_TRY(launch_cfft_lines((int)n2, p->tmp, dst, nb * n1, n2, 1, 1, n1, p->tw2, false, inverse, scale, nullptr,
                                      nullptr, 1, stream, 0, 0, n1, n, n));
        }
        return AFB_OK;
    }
    if (p->mode == 3) return launch_mixed_cfft(n, p->tw, in, out, batch, inverse, scale, stream);
    if (p->mode == 4)  // the element-wise halves of a split plan live in the callers' pre / post kernels
        return fail(AFB_UNSUPPORTED, "split FFT plans run through launch_*_pre / cfft_exec_work / launch_*_post");
    // Bluestein: X_k = chirp_k * sum_j (x_j chirp_j) conj(chirp)_{k-j}
    if (p->M <= LINE_MAX_N) {  // one kernel; in place is safe (a CTA reads its whole lines before it writes them)
        BlueParams bp{in, out, n, batch, p->chirp, p->bhat, p->sub->tw, inverse ? 1 : 0, inverse ? 1 : 0,
                      scale / (double)p->M};
        return launch_blue_fused(p->M, bp, stream);
    }
    if (p->bS > 1) {
        const int64_t h = LINE_MAX_N, M = p->M;
        for (int64_t b0 = 0; b0 < batch; b0 += p->max_batch) {
            const int64_t nb = std::min(p->max_batch, batch - b0);
            const unsigned g1 = ew_grid(nb * h), g2 = ew_grid(nb * n);
            const int ci = inverse ? 1 : 0;
#define AFB_BSPLIT(S_) \
    case S_: { auto k1 = blue_split_pre_kernel<S_>; \
               AFB_LAUNCH(k1, g1, 256, 0, stream, in + b0 * n, p->pad, n, h, nb, p->chirp, p->rootM, ci); break; }
            switch (p->bS) { AFB_BSPLIT(2) AFB_BSPLIT(4) AFB_BSPLIT(8) AFB_BSPLIT(16) default: return fail(AFB_BAD_ARG, "split factor"); }
#undef AFB_BSPLIT
            AFB_KERNEL_CHECK();
            BlueParams bp{p->pad, p->pad, h, nb * p->bS, nullptr, p->bhat, p->sub->tw, 0, 0, 1.0, p->bS};
            AFB_TRY(launch_blue_fused(h, bp, stream));
#define AFB_BSPLIT(S_) \
    case S_: { auto k2 = blue_split_post_kernel<S_>; \
               AFB_LAUNCH(k2, g2, 256, 0, stream, p->pad, out + b0 * n, n, h, nb, p->chirp, p->rootM, ci, scale / (double)M); break; }
            switch (p->bS) { AFB_BSPLIT(2) AFB_BSPLIT(4) AFB_BSPLIT(8) AFB_BSPLIT(16) default: return fail(AFB_BAD_ARG, "split factor"); }
#undef AFB_BSPLIT
            AFB_KERNEL_CHECK();
        }
        return AFB_OK;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += p->max_batch) {
        const int64_t nb = std::min(p->max_batch, batch - b0);
        const int64_t M = p->M;
        auto kpre = blue_pre_kernel;
        AFB_LAUNCH(kpre, ew_grid(nb * M), 256, 0, stream, in + b0 * n, p->pad, n, M, nb, p->chirp, inverse ? 1 : 0);
        AFB_KERNEL_CHECK();
        AFB_TRY(cfft_exec(p->sub, p->pad, p->pad, nb, false, 1.0, stream));
        auto kmul = blue_mul_kernel;
        AFB_LAUNCH(kmul, ew_grid(nb * M), 256, 0, stream, p->pad, M, nb, p->bhat);
        AFB_KERNEL_CHECK();
        AFB_TRY(cfft_exec(p->sub, p->pad, p->pad, nb, true, 1.0 / (double)M, stream));
        auto kpost = blue_post_kernel;
        AFB_LAUNCH(kpost, ew_grid(nb * n), 256, 0, stream, p->pad, out + b0 * n, n, M, nb, p->chirp, inverse ? 1 : 0, scale);
        AFB_KERNEL_CHECK();
    }
    return AFB_OK;
}

// =====================================================================================================
// generic pre / post kernels: real transform <-> length-n complex FFT (one thread per element)
//   qtab[k] = exp(-i pi k / (2n)), k = 0..n   (Chebyshev kinds only)
// =====================================================================================================
// The four REAL kinds transform two lines per complex FFT: work line q = z(line 2q) + i z(line 2q+1).  Inverse kinds
// produce real sequences, so the two results are the real and imaginary parts; forward kinds recover the spectra of the
// two real inputs from Z[k] and conj(Z[n-k]).
__device__ __forceinline__ bool generic_real_kind(int kind) { return kind <= AFB_FOURIER_INV; }

__device__ __forceinline__ cplx generic_pre_value(int kind, int64_t n, int64_t nh, const void* __restrict__ line, int64_t j,
                                                  const cplx* __restrict__ qtab) {
    cplx z = make_double2(0.0, 0.0);
    if (kind == AFB_CHEB1_FWD) {
        const double* v = reinterpret_cast<const double*>(line);
        z.x = (j < nh) ? v[2 * j] : v[2 * (n - 1 - j) + 1];
    } else if (kind == AFB_CHEB1_INV) {
        // X'_k = conj(w^k) (c'_k - i c'_{n-k}) / 2,  c'_0 = 2 c_0, c'_n = 0
        const double* c = reinterpret_cast<const double*>(line);
        if (j == 0) z.x = c[0];
        else {
            const cplx w = qtab[j];
            z = cmul(make_double2(0.5 * c[j], -0.5 * c[n - j]), make_double2(w.x, -w.y));
        }
    } else if (kind == AFB_FOURIER_FWD) {
        z.x = reinterpret_cast<const double*>(line)[j];
    } else if (kind == AFB_FOURIER_INV) {
        // F'_0 = c_0 ; F'_k = (a_k - i b_k)/2 (1 <= k < nh) ; F'_{n-k} = conj ; even n: F'_{n/2} = -c_last
        const double* c = reinterpret_cast<const double*>(line);
        if (j == 0) z.x = c[0];
        else if ((n % 2 == 0) && j == n / 2) z.x = -c[n - 1];
        else if (j < nh) z = make_double2(0.5 * c[2 * j], -0.5 * c[2 * j - 1]);
        else { const int64_t k = n - j; z = make_double2(0.5 * c[2 * k], 0.5 * c[2 * k - 1]); }
    } else if (kind == AFB_LAURENT_FWD || kind == AFB_TAYLOR_FWD || kind == AFB_HARDYM_FWD || kind == AFB_TAYLOR_INV) {
        z = reinterpret_cast<const cplx*>(line)[j];
    } else if (kind == AFB_HARDYM_INV) {  // F[n-1-k] = c[k]
        z = reinterpret_cast<const cplx*>(line)[n - 1 - j];
    } else {  // AFB_LAURENT_INV: F[m] = c[imap(m)]
        const int64_t i = (j < nh) ? 2 * j : 2 * (n - j) - 1;
        z = reinterpret_cast<const cplx*>(line)[i];
    }
    return z;
}

// element j of work line q: z(line 2q) + i z(line 2q + 1) for the real kinds, the line itself for the complex ones
__device__ __forceinline__ cplx generic_packed_value(int kind, bool real, int64_t n, int64_t nh, const void* __restrict__ in,
                                                     int64_t istride, int64_t batch, int64_t q, int64_t j,
                                                     const cplx* __restrict__ qtab) {
    if (!real) return generic_pre_value(kind, n, nh, reinterpret_cast<const cplx*>(in) + q * istride, j, qtab);
    const double* base = reinterpret_cast<const double*>(in);
    cplx z = generic_pre_value(kind, n, nh, base + (2 * q) * istride, j, qtab);
    if (2 * q + 1 < batch) {
        const cplx z2 = generic_pre_value(kind, n, nh, base + (2 * q + 1) * istride, j, qtab);
        z = make_double2(z.x - z2.y, z.y + z2.x);  // z + i z2
    }
    return z;
}

__global__ void generic_pre_kernel(int kind, int64_t n, const void* __restrict__ in, int64_t istride,
                                   cplx* __restrict__ work, int64_t batch, const cplx* __restrict__ qtab) {
    const bool real = generic_real_kind(kind);
    const int64_t total = (real ? (batch + 1) / 2 : batch) * n;
    const int64_t nh = (n + 1) / 2;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x)
        work[g] = generic_packed_value(kind, real, n, nh, in, istride, batch, g / n, g % n, qtab);
}

// ---- split plans (mode 4): n = S h.  One radix-S decimation-in-frequency pass
//        y_s[j] = w_n^(+-j s) sum_r z[j + r h] w_S^(+-r s),   j < h, s < S,
// fused into the pre kernels, leaves S contiguous blocks of length h per line; their length-h transforms (one-CTA engines:
// mixed radix or fused Bluestein) hold X[S k + s] = Y_s[k], which the post kernels read through SpecView.
template <int S>
__device__ __forceinline__ void split_dif(const cplx (&z)[S], int64_t j, int64_t h, const cplx* __restrict__ rootn, bool inv,
                                          cplx* __restrict__ dst) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
        cplx acc = z[0];
#pragma unroll
        for (int r = 1; r < S; ++r) {
            cplx w = rootn[(int64_t)((r * s) % S) * h];
            if (inv) w.y = -w.y;
            acc = cadd(acc, cmul(z[r], w));
        }
        if (s) {
            cplx w = rootn[j * s];
            if (inv) w.y = -w.y;
            acc = cmul(acc, w);
        }
        dst[(int64_t)s * h + j] = acc;
    }
}
struct SpecView {   // element k of a spectrum left as S blocks of length h
    const cplx* base;
    int64_t h;
    int S;
    __device__ __forceinline__ cplx operator[](int64_t k) const { return S == 1 ? base[k] : base[(k % S) * h + k / S]; }
};

template <int S>
__global__ void generic_pre_split_kernel(int kind, int64_t n, const void* __restrict__ in, int64_t istride,
                                         cplx* __restrict__ work, int64_t batch, const cplx* __restrict__ qtab,
                                         const cplx* __restrict__ rootn, int inv) {
    const bool real = generic_real_kind(kind);
    const int64_t h = n / S, total = (real ? (batch + 1) / 2 : batch) * h, nh = (n + 1) / 2;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = g / h, j = g % h;
        cplx z[S];
#pragma unroll
        for (int r = 0; r < S; ++r) z[r] = generic_packed_value(kind, real, n, nh, in, istride, batch, q, j + r * h, qtab);
        split_dif<S>(z, j, h, rootn, inv != 0, work + q * n);
    }
}

// ---- Bluestein with a padded length M = S h above one CTA (h = 8192): the radix-S DIF pass of FFT_M rides in the chirp
// kernel, the blocks go through blue_fused_kernel<h> as plain circular convolutions (FFT_h, filter block, inverse FFT_h) and
// the radix-S DIT pass of the inverse FFT_M rides in the final chirp kernel: three launches, 4 passes over the padded data
// instead of 12.
template <int S>
__global__ void blue_split_pre_kernel(const cplx* __restrict__ in, cplx* __restrict__ pad, int64_t n, int64_t h, int64_t batch,
                                      const cplx* __restrict__ chirp, const cplx* __restrict__ rootM, int conj_in) {
    const int64_t total = batch * h;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / h, j = g % h;
        cplx z[S];
#pragma unroll
        for (int r = 0; r < S; ++r) {
            const int64_t i = j + r * h;
            z[r] = make_double2(0.0, 0.0);
            if (i < n) {
                cplx v = in[b * n + i];
                if (conj_in) v.y = -v.y;
                z[r] = cmul(v, chirp[i]);
            }
        }
        split_dif<S>(z, j, h, rootM, false, pad + b * (S * h));
    }
}
template <int S>
__global__ void blue_split_post_kernel(const cplx* __restrict__ pad, cplx* __restrict__ out, int64_t n, int64_t h, int64_t batch,
                                       const cplx* __restrict__ chirp, const cplx* __restrict__ rootM, int conj_out, double scale) {
    const int64_t total = batch * n, M = S * h;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / n, k = g % n;
        const cplx* d = pad + b * M + (k % h);
        cplx acc = d[0];
#pragma unroll
        for (int s = 1; s < S; ++s) {
            const cplx w = rootM[(s * k) & (M - 1)];
            acc = cadd(acc, cmul(d[(int64_t)s * h], make_double2(w.x, -w.y)));
        }
        cplx z = cscale(cmul(acc, chirp[k]), scale);
        if (conj_out) z.y = -z.y;
        out[g] = z;
    }
}

// post kernel of the COMPLEX kinds (the real kinds: generic_post_pair_kernel below)
__global__ void generic_post_kernel(int kind, int64_t n, const cplx* __restrict__ work, void* __restrict__ out,
                                    int64_t ostride, int64_t batch, const cplx* __restrict__ qtab, int S) {
    (void)qtab;
    const int64_t total = batch * n;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / n, o = g % n;
        const SpecView W{work + b * n, n / S, S};
        cplx* dst = reinterpret_cast<cplx*>(out) + b * ostride;
        if (kind == AFB_LAURENT_FWD) {
            const int64_t m = (o & 1) ? n - ((o + 1) >> 1) : (o >> 1);
            dst[o] = cscale(W[m], 1.0 / (double)n);
        } else if (kind == AFB_TAYLOR_FWD) {        // f_k = F_k / n
            dst[o] = cscale(W[o], 1.0 / (double)n);
        } else if (kind == AFB_HARDYM_FWD) {        // f_k = F_{n-1-k} / n   (mode -(k+1))
            dst[o] = cscale(W[n - 1 - o], 1.0 / (double)n);
        } else {
            dst[o] = W[o];
        }
    }
}

// output o of the two real lines packed in one complex transform: W(k) = element k of the length-n (inverse) DFT of
// z(line 2q) + i z(line 2q + 1); ve belongs to line 2q, vo to line 2q + 1
template <class WF>
__device__ __forceinline__ void real_pair_post(int kind, int64_t n, int64_t o, const WF& W, const cplx* __restrict__ qtab,
                                               double& ve, double& vo) {
    const double dn = (double)n;
    if (kind == AFB_CHEB1_INV) {
        const cplx z = W((o & 1) ? n - 1 - (o - 1) / 2 : o / 2);
        ve = z.x; vo = z.y;
    } else if (kind == AFB_FOURIER_INV) {
        const cplx z = W(o);
        ve = z.x; vo = z.y;
    } else {
        const bool nyq = (kind == AFB_FOURIER_FWD) && (n % 2 == 0) && o == n - 1;
        const int64_t k = (kind == AFB_CHEB1_FWD) ? o : (o == 0 ? 0 : nyq ? n / 2 : (o & 1) ? (o + 1) / 2 : o / 2);
        const cplx A = W(k), Bc = W((n - k) % n);
        const cplx Ye = make_double2(0.5 * (A.x + Bc.x), 0.5 * (A.y - Bc.y));    // (A + conj B) / 2
        const cplx Yo = make_double2(0.5 * (A.y + Bc.y), -0.5 * (A.x - Bc.x));   // (A - conj B) / (2i)
        if (kind == AFB_CHEB1_FWD) {
            const cplx w = ldg2(qtab + o);
            const double sc = (o == 0) ? 1.0 / dn : 2.0 / dn;
            ve = cmul(Ye, w).x * sc; vo = cmul(Yo, w).x * sc;
        } else if (o == 0) { ve = Ye.x / dn; vo = Yo.x / dn; }
        else if (nyq) { ve = -Ye.x / dn; vo = -Yo.x / dn; }
        else if (o & 1) { ve = -Ye.y * (2.0 / dn); vo = -Yo.y * (2.0 / dn); }
        else { ve = Ye.x * (2.0 / dn); vo = Yo.x * (2.0 / dn); }
    }
}

// the real kinds of the multi-kernel route: one thread per (pair of lines, output index) -- every spectrum element is read
// for both lines at once
__global__ void generic_post_pair_kernel(int kind, int64_t n, const cplx* __restrict__ work, double* __restrict__ out,
                                         int64_t ostride, int64_t batch, const cplx* __restrict__ qtab, int S) {
    const int64_t total = ((batch + 1) / 2) * n;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = g / n, o = g % n;
        const SpecView V{work + q * n, n / S, S};
        auto W = [&](int64_t k) -> cplx { return V[k]; };
        double ve, vo;
        real_pair_post(kind, n, o, W, qtab, ve, vo);
        out[(2 * q) * ostride + o] = ve;
        if (2 * q + 1 < batch) out[(2 * q + 1) * ostride + o] = vo;
    }
}

// ---- the four real kinds of arbitrary length in ONE kernel (padded Bluestein length M <= 8192) -------------------
// pre-processing of two real lines in the load loop, the two shared-memory FFTs of blue_fused_kernel, then the
// spectrum split and the kind's post-processing straight from shared memory: one read and one write of the data.
struct BlueRealParams {
    const double* in;
    double* out;
    int64_t n, batch, istride, ostride;  // batch = real lines
    const cplx *chirp, *bhat, *tw, *qtab;
    double scale;  // 1/M
};
template <int M, int KIND>
__global__ void __launch_bounds__(BlueGeom<M>::CT, M == 8192 ? 1 : (M == 4096 && AFB_R32_4096 ? 2 : 512 / BlueGeom<M>::CT)) blue_real_kernel(const __grid_constant__ BlueRealParams p) {
    typedef FftCfg<M> C;
    typedef BlueGeom<M> G;
    constexpr bool INV = (KIND & 1) != 0;
    AFB_DYN_SMEM(cplx, smem);
    const int l = threadIdx.x / C::T, t = threadIdx.x % C::T;
    const int64_t q = (int64_t)blockIdx.x * G::LPC + l;  // pair of lines 2q, 2q+1
    const bool act = 2 * q < p.batch;
    const int64_t qc = act ? q : 0;
    const bool has2 = 2 * qc + 1 < p.batch;
    const int64_t n = p.n, nh = (n + 1) / 2;
    cplx* s = smem + (size_t)l * FftLine<M>::STRIDE;
    const double* l0 = p.in + (2 * qc) * p.istride;
    const double* l1 = has2 ? l0 + p.istride : l0;
    cplx x[C::E];
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int64_t j = t + C::T * r;
        const int64_t jc = j < n ? j : n - 1;
        const cplx z1 = generic_pre_value(KIND, n, nh, l0, jc, p.qtab);
        cplx z2 = generic_pre_value(KIND, n, nh, l1, jc, p.qtab);
        if (!has2) z2 = make_double2(0.0, 0.0);
        cplx z = make_double2(z1.x - z2.y, z1.y + z2.x);  // z1 + i z2
        if (INV) z.y = -z.y;
        z = cmul(z, ldg2(p.chirp + jc));
        x[r] = (j < n) ? z : make_double2(0.0, 0.0);
    }
    fft_line<M>(s, t, x, p.tw);
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int i = t + C::T * r;
        const cplx z = cmul(line_get<M>(s, i), ldg2(p.bhat + i));
        x[r] = make_double2(z.x, -z.y);
    }
    __syncthreads();
    int t2 = t;  // see blue_fused_kernel
    const cplx* tw2 = p.tw;
    cplx* s2 = s;
#ifndef AFB_HOST_EMU
    asm volatile("" : "+r"(t2), "+l"(tw2), "+l"(s2));
#endif
    fft_line<M>(s2, t2, x, tw2);
    if (!act) return;
    // W(k): element k of the length-n (inverse) DFT of the packed line
    auto W = [&](int64_t k) -> cplx {
        const cplx y = line_get<M>(s2, (int)k);
        cplx z = cscale(cmul(make_double2(y.x, -y.y), ldg2(p.chirp + k)), p.scale);
        if (INV) z.y = -z.y;
        return z;
    };
    double* o0 = p.out + (2 * q) * p.ostride;
    double* o1 = o0 + p.ostride;
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int64_t o = t2 + C::T * r;
        if (o >= n) continue;
        double ve, vo;
        real_pair_post(KIND, n, o, W, p.qtab, ve, vo);
        o0[o] = ve;
        if (has2) o1[o] = vo;
    }
}
template <int M, int KIND> static int32_t launch_blue_real_mk(const BlueRealParams& p, void* stream) {
    typedef BlueGeom<M> G;
    auto k = blue_real_kernel<M, KIND>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    const int64_t npairs = (p.batch + 1) / 2;
    AFB_LAUNCH(k, (unsigned)((npairs + G::LPC - 1) / G::LPC), G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
template <int KIND> static int32_t launch_blue_real_k(int64_t M, const BlueRealParams& p, void* stream) {
    switch (M) {
        case 32: return launch_blue_real_mk<32, KIND>(p, stream);
        case 64: return launch_blue_real_mk<64, KIND>(p, stream);
        case 128: return launch_blue_real_mk<128, KIND>(p, stream);
        case 256: return launch_blue_real_mk<256, KIND>(p, stream);
        case 512: return launch_blue_real_mk<512, KIND>(p, stream);
        case 1024: return launch_blue_real_mk<1024, KIND>(p, stream);
        case 2048: return launch_blue_real_mk<2048, KIND>(p, stream);
        case 4096: return launch_blue_real_mk<4096, KIND>(p, stream);
        case 8192: return launch_blue_real_mk<8192, KIND>(p, stream);
        default: return fail(AFB_UNSUPPORTED, "fused Bluestein: padded length not instantiated");
    }
}
// true when plan p is a Bluestein plan whose padded length fits one CTA: the real kinds then run in a single kernel
// ... or a mixed-radix plan (one length-n FFT in shared memory, afb_mixed.cu)
bool cfft_real_fused_ok(const CfftPlan* p) { return p && ((p->mode == 2 && p->M <= LINE_MAX_N) || p->mode == 3); }
int32_t launch_blue_real(CfftPlan* p, int kind, const double* in, int64_t istride, double* out, int64_t ostride, int64_t batch,
                         const cplx* qtab, void* stream) {
    if (batch == 0) return AFB_OK;
    if (p->mode == 3) return launch_mixed_real(kind, p->n, p->tw, qtab, in, istride, out, ostride, batch, stream);
    BlueRealParams bp{in, out, p->n, batch, istride, ostride, p->chirp, p->bhat, p->sub->tw, qtab, 1.0 / (double)p->M};
    switch (kind) {
        case AFB_CHEB1_FWD: return launch_blue_real_k<AFB_CHEB1_FWD>(p->M, bp, stream);
        case AFB_CHEB1_INV: return launch_blue_real_k<AFB_CHEB1_INV>(p->M, bp, stream);
        case AFB_FOURIER_FWD: return launch_blue_real_k<AFB_FOURIER_FWD>(p->M, bp, stream);
        case AFB_FOURIER_INV: return launch_blue_real_k<AFB_FOURIER_INV>(p->M, bp, stream);
        default: return fail(AFB_BAD_ARG, "fused real Bluestein: not a real kind");
    }
}

// =====================================================================================================
// Extension kinds: transforms defined through the FFT of an even / odd extension (complex length cn):
//   AFB_CHEB2_FWD/INV  second-kind points, FFTW REDFT00 (UPSTREAM FastTransforms plan_(i)chebyshevtransform(x, Val(2)));
//                      even extension of length cn = 2(n-1):  Y_k = Re FFT([x_0..x_{n-1}, x_{n-2}..x_1])_k
//                      fwd: c = Y/(n-1), c_0 /= 2, c_{n-1} /= 2 ;  inv: c_0 *= 2, c_{n-1} *= 2, v = Y/2
//   AFB_SIN_FWD/INV    SinSpace, /root/reference/src/Extras/fftGeneric.jl:74-81, restated verbatim:
//                      v = imag(fft([0; -x; 0; reverse(x)]))[2:n+1],  fwd: v/(n+1),  inv: v/2     (cn = 2n+2)
// =====================================================================================================
// The extensions are real with a real (even) or purely imaginary (odd) spectrum, so TWO lines share one complex FFT:
// work line q holds ext(line 2q) + i ext(line 2q+1);  FFT = E_{2q} + i E_{2q+1}  (even)  /  -S_{2q+1} + i S_{2q}  (odd, E = iS).
__device__ __forceinline__ double ext_value(int kind, int64_t n, int64_t cn, const double* __restrict__ x, int64_t j) {
    if (kind == AFB_CHEB2_FWD || kind == AFB_CHEB2_INV) {
        const int64_t i = (j < n) ? j : cn - j;
        const double v = x[i];
        return (kind == AFB_CHEB2_INV && (i == 0 || i == n - 1)) ? v * 2.0 : v;
    }
    // [0; -x; 0; reverse(x)]
    if (j == 0 || j == n + 1) return 0.0;
    return (j <= n) ? -x[j - 1] : x[2 * n + 1 - j];
}
__global__ void ext_pre_kernel(int kind, int64_t n, int64_t cn, const double* __restrict__ in, int64_t istride,
                               cplx* __restrict__ work, int64_t batch) {
    const int64_t total = ((batch + 1) / 2) * cn;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = g / cn, j = g % cn;
        const double re = ext_value(kind, n, cn, in + (2 * q) * istride, j);
        const double im = (2 * q + 1 < batch) ? ext_value(kind, n, cn, in + (2 * q + 1) * istride, j) : 0.0;
        work[g] = make_double2(re, im);
    }
}
template <int S>
__global__ void ext_pre_split_kernel(int kind, int64_t n, int64_t cn, const double* __restrict__ in, int64_t istride,
                                     cplx* __restrict__ work, int64_t batch, const cplx* __restrict__ rootn) {
    const int64_t h = cn / S, total = ((batch + 1) / 2) * h;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = g / h, j = g % h;
        const bool has2 = 2 * q + 1 < batch;
        cplx z[S];
#pragma unroll
        for (int r = 0; r < S; ++r) {
            z[r].x = ext_value(kind, n, cn, in + (2 * q) * istride, j + r * h);
            z[r].y = has2 ? ext_value(kind, n, cn, in + (2 * q + 1) * istride, j + r * h) : 0.0;
        }
        split_dif<S>(z, j, h, rootn, false, work + q * cn);
    }
}
__global__ void ext_post_kernel(int kind, int64_t n, int64_t cn, const cplx* __restrict__ work, double* __restrict__ out,
                                int64_t ostride, int64_t batch, int S) {
    const int64_t total = batch * n;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / n, k = g % n;
        const SpecView W{work + (b >> 1) * cn, cn / S, S};
        const bool odd = b & 1;
        double r;
        if (kind == AFB_CHEB2_FWD) {
            r = (odd ? W[k].y : W[k].x) / (double)(n - 1);
            if (k == 0 || k == n - 1) r /= 2.0;
        } else if (kind == AFB_CHEB2_INV) {
            r = (odd ? W[k].y : W[k].x) / 2.0;
        } else {
            const double sv = odd ? -W[k + 1].x : W[k + 1].y;
            r = (kind == AFB_SIN_FWD) ? sv / (double)(n + 1) : sv / 2.0;
        }
        out[b * ostride + k] = r;
    }
}
// ---- extension kinds in ONE kernel when the FFT (cn a power of two <= 8192, or Bluestein with M <= 8192) fits a CTA:
// the extension of two real lines is formed in the load loop, the transform(s) run in shared memory and the wanted
// real / imaginary parts are scaled and stored straight from there.
struct ExtParams {
    const double* in;
    double* out;
    int64_t n, cn, batch, istride, ostride;  // batch = real lines
    const cplx *chirp, *bhat, *tw;
    double scale;  // 1/M (Bluestein)
};
template <int M, int KIND, bool BLUE>
__global__ void __launch_bounds__(BlueGeom<M>::CT, M == 8192 ? 1 : (M == 4096 && AFB_R32_4096 ? 2 : 512 / BlueGeom<M>::CT)) ext_fused_kernel(const __grid_constant__ ExtParams p) {
    typedef FftCfg<M> C;
    typedef BlueGeom<M> G;
    AFB_DYN_SMEM(cplx, smem);
    const int l = threadIdx.x / C::T, t = threadIdx.x % C::T;
    const int64_t q = (int64_t)blockIdx.x * G::LPC + l;
    const bool act = 2 * q < p.batch;
    const int64_t qc = act ? q : 0;
    const bool has2 = 2 * qc + 1 < p.batch;
    const int64_t n = p.n, cn = p.cn;
    cplx* s = smem + (size_t)l * FftLine<M>::STRIDE;
    const double* l0 = p.in + (2 * qc) * p.istride;
    const double* l1 = has2 ? l0 + p.istride : l0;
    cplx x[C::E];
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int64_t j = t + C::T * r;
        const int64_t jc = j < cn ? j : cn - 1;
        cplx z = make_double2(ext_value(KIND, n, cn, l0, jc), ext_value(KIND, n, cn, l1, jc));
        if (!has2) z.y = 0.0;
        if (BLUE) z = cmul(z, ldg2(p.chirp + jc));
        x[r] = (j < cn) ? z : make_double2(0.0, 0.0);
    }
    fft_line<M>(s, t, x, p.tw);
    int t2 = t;
    cplx* s2 = s;
    if (BLUE) {
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int i = t + C::T * r;
            const cplx z = cmul(line_get<M>(s, i), ldg2(p.bhat + i));
            x[r] = make_double2(z.x, -z.y);
        }
        __syncthreads();
        const cplx* tw2 = p.tw;  // see blue_fused_kernel
#ifndef AFB_HOST_EMU
        asm volatile("" : "+r"(t2), "+l"(tw2), "+l"(s2));
#endif
        fft_line<M>(s2, t2, x, tw2);
    }
    if (!act) return;
    auto W = [&](int64_t k) -> cplx {
        const cplx y = line_get<M>(s2, (int)k);
        if (!BLUE) return y;
        return cscale(cmul(make_double2(y.x, -y.y), ldg2(p.chirp + k)), p.scale);
    };
    double* o0 = p.out + (2 * q) * p.ostride;
    double* o1 = o0 + p.ostride;
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int64_t k = t2 + C::T * r;
        if (k >= n) continue;
        double ve, vo;
        if (KIND == AFB_CHEB2_FWD || KIND == AFB_CHEB2_INV) {
            const cplx w = W(k);
            double sc = (KIND == AFB_CHEB2_FWD) ? 1.0 / (double)(n - 1) : 0.5;
            if (KIND == AFB_CHEB2_FWD && (k == 0 || k == n - 1)) sc *= 0.5;
            ve = w.x * sc; vo = w.y * sc;
        } else {
            const cplx w = W(k + 1);
            const double sc = (KIND == AFB_SIN_FWD) ? 1.0 / (double)(n + 1) : 0.5;
            ve = w.y * sc; vo = -w.x * sc;
        }
        o0[k] = ve;
        if (has2) o1[k] = vo;
    }
}
template <int M, int KIND, bool BLUE> static int32_t launch_ext_fused_mk(const ExtParams& p, void* stream) {
    typedef BlueGeom<M> G;
    auto k = ext_fused_kernel<M, KIND, BLUE>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    const int64_t npairs = (p.batch + 1) / 2;
    AFB_LAUNCH(k, (unsigned)((npairs + G::LPC - 1) / G::LPC), G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
template <int KIND, bool BLUE> static int32_t launch_ext_fused_k(int64_t M, const ExtParams& p, void* stream) {
    switch (M) {
        case 32: return launch_ext_fused_mk<32, KIND, BLUE>(p, stream);
        case 64: return launch_ext_fused_mk<64, KIND, BLUE>(p, stream);
        case 128: return launch_ext_fused_mk<128, KIND, BLUE>(p, stream);
        case 256: return launch_ext_fused_mk<256, KIND, BLUE>(p, stream);
        case 512: return launch_ext_fused_mk<512, KIND, BLUE>(p, stream);
        case 1024: return launch_ext_fused_mk<1024, KIND, BLUE>(p, stream);
        case 2048: return launch_ext_fused_mk<2048, KIND, BLUE>(p, stream);
        case 4096: return launch_ext_fused_mk<4096, KIND, BLUE>(p, stream);
        case 8192: return launch_ext_fused_mk<8192, KIND, BLUE>(p, stream);
        default: return fail(AFB_UNSUPPORTED, "fused extension transform: length not instantiated");
    }
}
bool cfft_ext_fused_ok(const CfftPlan* p) {
    return p && ((p->mode == 0) || (p->mode == 2 && p->M <= LINE_MAX_N));
}
int32_t launch_ext_fused(CfftPlan* p, int kind, int64_t n, const double* in, int64_t istride, double* out, int64_t ostride,
                         int64_t batch, void* stream) {
    if (batch == 0) return AFB_OK;
    const bool blue = p->mode == 2;
    const int64_t M = blue ? p->M : p->n;
    ExtParams ep{in, out, n, p->n, batch, istride, ostride, p->chirp, p->bhat, blue ? p->sub->tw : p->tw, 1.0 / (double)M};
#define AFB_EXT_CASE(K_) \
    case K_: return blue ? launch_ext_fused_k<K_, true>(M, ep, stream) : launch_ext_fused_k<K_, false>(M, ep, stream)
    switch (kind) {
        AFB_EXT_CASE(AFB_CHEB2_FWD);
        AFB_EXT_CASE(AFB_CHEB2_INV);
        AFB_EXT_CASE(AFB_SIN_FWD);
        AFB_EXT_CASE(AFB_SIN_INV);
        default: return fail(AFB_BAD_ARG, "fused extension transform: not an extension kind");
    }
#undef AFB_EXT_CASE
}

int32_t launch_ext_pre(int kind, int64_t n, int64_t cn, const double* in, int64_t istride, cplx* work, int64_t batch,
                       void* stream, const CfftPlan* sp) {
    if (batch == 0) return AFB_OK;
    const int S = cfft_split(sp);
    if (S > 1) {
        const unsigned grid = ew_grid(((batch + 1) / 2) * (cn / S));
#define AFB_EXT_SPLIT(S_) \
    case S_: { auto k = ext_pre_split_kernel<S_>; \
               AFB_LAUNCH(k, grid, 256, 0, stream, kind, n, cn, in, istride, work, batch, sp->rootn); break; }
        switch (S) {
            AFB_EXT_SPLIT(2) AFB_EXT_SPLIT(3) AFB_EXT_SPLIT(4) AFB_EXT_SPLIT(5) AFB_EXT_SPLIT(7) AFB_EXT_SPLIT(8) AFB_EXT_SPLIT(16)
            default: return fail(AFB_BAD_ARG, "split factor");
        }
#undef AFB_EXT_SPLIT
        AFB_KERNEL_CHECK();
        return AFB_OK;
    }
    auto k = ext_pre_kernel;
    AFB_LAUNCH(k, ew_grid(((batch + 1) / 2) * cn), 256, 0, stream, kind, n, cn, in, istride, work, batch);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_ext_post(int kind, int64_t n, int64_t cn, const cplx* work, double* out, int64_t ostride, int64_t batch,
                        void* stream, const CfftPlan* sp) {
    if (batch == 0) return AFB_OK;
    auto k = ext_post_kernel;
    AFB_LAUNCH(k, ew_grid(batch * n), 256, 0, stream, kind, n, cn, work, out, ostride, batch, cfft_split(sp));
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

int32_t launch_generic_pre(int kind, int64_t n, const void* in, int64_t istride, cplx* work, int64_t batch,
                           const cplx* qtab, void* stream, const CfftPlan* sp) {
    if (batch == 0) return AFB_OK;
    const int S = cfft_split(sp);
    if (S > 1) {
        const unsigned grid = ew_grid(batch * (n / S));
        const int inv = kind & 1;   // inverse kinds ask cfft_exec for the inverse transform
#define AFB_GEN_SPLIT(S_) \
    case S_: { auto k = generic_pre_split_kernel<S_>; \
               AFB_LAUNCH(k, grid, 256, 0, stream, kind, n, in, istride, work, batch, qtab, sp->rootn, inv); break; }
        switch (S) {
            AFB_GEN_SPLIT(2) AFB_GEN_SPLIT(3) AFB_GEN_SPLIT(4) AFB_GEN_SPLIT(5) AFB_GEN_SPLIT(7) AFB_GEN_SPLIT(8) AFB_GEN_SPLIT(16)
            default: return fail(AFB_BAD_ARG, "split factor");
        }
#undef AFB_GEN_SPLIT
        AFB_KERNEL_CHECK();
        return AFB_OK;
    }
    auto k = generic_pre_kernel;
    AFB_LAUNCH(k, ew_grid(batch * n), 256, 0, stream, kind, n, in, istride, work, batch, qtab);  // grid-stride: covers the packed count too
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_generic_post(int kind, int64_t n, const cplx* work, void* out, int64_t ostride, int64_t batch,
                            const cplx* qtab, void* stream, const CfftPlan* sp) {
    if (batch == 0) return AFB_OK;
    if (kind <= AFB_FOURIER_INV) {
        auto kp = generic_post_pair_kernel;
        AFB_LAUNCH(kp, ew_grid(((batch + 1) / 2) * n), 256, 0, stream, kind, n, work, reinterpret_cast<double*>(out), ostride, batch,
                   qtab, cfft_split(sp));
        AFB_KERNEL_CHECK();
        return AFB_OK;
    }
    auto k = generic_post_kernel;
    AFB_LAUNCH(k, ew_grid(batch * n), 256, 0, stream, kind, n, work, out, ostride, batch, qtab, cfft_split(sp));
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// =====================================================================================================
// Large power-of-two real transforms (n >= 32768; config 1 is n = 2^20): ONE half-length complex four-step FFT
// (N = n/2 = n1 n2) whose step A gathers straight from the real array in DCT order (or whose step B scatters
// back to it), plus one pair kernel doing the real-FFT split / quarter-wave twiddle with sincospi twiddles.
// Three launches and 6 x 8N bytes of (L2-resident) traffic per transform instead of four launches and 14 x 8N.
// =====================================================================================================
int32_t cfft_exec_dctorder(CfftPlan* p, const void* in, void* out, bool conj_in, bool conj_out, double scale, int io_mode,
                           void* stream) {
    if (p->mode != 1) return fail(AFB_BAD_ARG, "cfft_exec_dctorder needs a four-step plan");
    const int64_t n = p->n, n1 = p->n1, n2 = p->n2;
    AFB_TRY(launch_cfft_lines((int)n1, reinterpret_cast<const cplx*>(in), p->tmp, n2, 1, n2, 1, n2, p->tw1, conj_in, false,
                              1.0, p->twA, p->twB, n2, stream, io_mode & 1, n));
    AFB_TRY(launch_cfft_lines((int)n2, p->tmp, reinterpret_cast<cplx*>(out), n1, n2, 1, 1, n1, p->tw2, false, conj_out, scale,
                              nullptr, nullptr, 1, stream, io_mode & 2, n));
    return AFB_OK;
}

__device__ __forceinline__ cplx unit_pi(double x) {  // exp(-i pi x)
    double sn, cs;
    sincospi(x, &sn, &cs);
    return make_double2(cs, -sn);
}

// forward: Z (length N = n/2, natural order) -> coefficients
__global__ void large_post_kernel(int kind, int64_t N, const cplx* __restrict__ Z, double* __restrict__ c) {
    const int64_t n = 2 * N;
    const double inv_n = 1.0 / (double)n, h = 0.70710678118654752440;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= N / 2; k += (int64_t)gridDim.x * blockDim.x) {
        if (k == 0) {
            const cplx A = Z[0];
            c[0] = (A.x + A.y) * inv_n;
            if (kind == AFB_CHEB1_FWD) c[N] = (A.x - A.y) * (inv_n * 1.41421356237309504880);
            else c[n - 1] = -(A.x - A.y) * inv_n;
            continue;
        }
        const cplx A = Z[k], B = Z[N - k];
        const cplx S = make_double2(A.x + B.x, A.y - B.y);
        const cplx D = make_double2(A.x - B.x, A.y + B.y);
        if (kind == AFB_CHEB1_FWD) {
            const cplx w1 = unit_pi((double)k / (double)(2 * n));
            const cplx w5 = unit_pi((double)(5 * k) / (double)(2 * n));
            const cplx U = cmul(S, make_double2(w1.x * inv_n, w1.y * inv_n));
            const cplx V = cmul(D, make_double2(w5.y * inv_n, -w5.x * inv_n));   // -i w^(5k) / n
            const cplx P = cadd(U, V), Q = csub(U, V);
            c[k] = P.x;
            c[n - k] = -P.y;
            if (2 * k != N) {
                c[N - k] = (Q.x - Q.y) * h;
                c[N + k] = (Q.x + Q.y) * h;
            }
        } else {  // AFB_FOURIER_FWD
            const cplx w = unit_pi((double)(2 * k) / (double)n);                     // exp(-2 pi i k / n)
            const cplx G = cmul(D, make_double2(w.y * inv_n, -w.x * inv_n));        // -i w / n
            const cplx Sn = make_double2(S.x * inv_n, S.y * inv_n);
            c[2 * k] = Sn.x + G.x;
            c[2 * k - 1] = -(Sn.y + G.y);
            if (2 * k != N) {
                c[2 * (N - k)] = Sn.x - G.x;
                c[2 * (N - k) - 1] = Sn.y - G.y;
            }
        }
    }
}

// inverse: coefficients -> conj(Z') (length N, natural order), 1/N folded in
__global__ void large_pre_kernel(int kind, int64_t N, const double* __restrict__ c, cplx* __restrict__ Zc) {
    const int64_t n = 2 * N;
    const double h = 0.70710678118654752440;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= N / 2; k += (int64_t)gridDim.x * blockDim.x) {
        if (k == 0) {
            if (kind == AFB_CHEB1_INV) Zc[0] = make_double2(c[0] + c[N] * h, -(c[0] - c[N] * h));
            else Zc[0] = make_double2(c[0] - c[n - 1], -(c[0] + c[n - 1]));
            continue;
        }
        cplx G, H;
        if (kind == AFB_CHEB1_INV) {
            const cplx w = unit_pi((double)k / (double)(2 * n));
            const cplx wk = make_double2(w.x, -w.y);                                  // conj(w^k)
            G = cmul(make_double2(c[k], -c[n - k]), wk);
            const cplx e = make_double2((wk.x + wk.y) * h, (wk.x - wk.y) * h);       // conj(w^(N-k))
            H = cmul(make_double2(c[N - k], -c[N + k]), e);
        } else {
            G = make_double2(c[2 * k], -c[2 * k - 1]);
            H = make_double2(c[2 * (N - k)], -c[2 * (N - k) - 1]);
        }
        const cplx t = unit_pi(-(double)(2 * k) / (double)n);                        // exp(+2 pi i k / n)
        const cplx Xe = make_double2(0.5 * (G.x + H.x), 0.5 * (G.y - H.y));
        const cplx W = make_double2(0.5 * (G.x - H.x), 0.5 * (G.y + H.y));
        const cplx Xo = cmul(W, t);
        Zc[k] = make_double2(Xe.x - Xo.y, -Xe.y - Xo.x);
        Zc[N - k] = make_double2(Xe.x + Xo.y, Xe.y - Xo.x);
    }
}

int32_t launch_large_post(int kind, int64_t n, const cplx* Z, double* out, void* stream) {
    auto k = large_post_kernel;
    AFB_LAUNCH(k, ew_grid(n / 4 + 1), 256, 0, stream, kind, n / 2, Z, out);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_large_pre(int kind, int64_t n, const double* in, cplx* Zc, void* stream) {
    auto k = large_pre_kernel;
    AFB_LAUNCH(k, ew_grid(n / 4 + 1), 256, 0, stream, kind, n / 2, in, Zc);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

}  // namespace afb
