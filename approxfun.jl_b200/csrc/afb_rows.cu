// Chebyshev transforms along the SECOND dimension of a column-major n1 x n2 matrix without transposing it
// (the row pass of transform!(::TensorSpace, M)).
//
// Two adjacent rows (2i', 2i'+1) are one complex sequence z_j = M[2i', j] + i M[2i'+1, j] -- in memory a
// complex matrix with leading dimension n1/2, contiguous along i'.  The DCT is real-linear, so the DCT of z
// is the DCT of both rows at once.  With Makhoul's permutation y[m] = z[2m], y[n-1-m] = z[2m+1] and
// Y = FFT_n(y), w = exp(-i pi / (2n)):
//     forward   c_k = (w^k Y[k] + conj(w^k) Y[(n-k) mod n]) / n          (k = 0: / 2n)
//     inverse   Y[k] = conj(w^k) (c'_k - i c'_{n-k}) / 2,  c'_0 = 2 c_0, c'_n = 0,   y = sum_k Y[k] e^{+2 pi i km/n}
// The length-n = NA*NB FFT over the strided direction is a four-step: m = m1 NB + m0, k = k1 + NA k0,
//     Y[k1 + NA k0] = sum_m0 w_n^(k1 m0) ( sum_m1 y[m1 NB + m0] w_NA^(m1 k1) ) w_NB^(m0 k0)
// run as two kernels, each a shared-memory FFT (afb_fft.cuh) on tiles of 16 / 8 adjacent row pairs so that
// every global access is a 256 / 128-byte contiguous run:
//     outer: fixed m0, FFT_NA over m1 (+ gather through the permutation, twiddle)      in  -> tmp[i', k1 NB + m0]
//     inner: groups k1 and NA-k1 together, FFT_NB over m0, then the k <-> n-k combine  tmp -> out[i', k]
// 3 passes over the matrix (48 B/element with the column pass) instead of 4 with explicit transposes.
#include "afb_fft.cuh"
#include "afb_internal.h"

namespace afb {

constexpr int INNER_CT = 128;      // inner kernel: 8 row pairs x 2 groups at NB = 128 (256 threads measured slower)
constexpr int64_t ROWS_PAD = 16;  // complex slots of padding per column of the intermediate

struct RowsParams {
    const cplx* in;
    cplx* out;
    int64_t half;     // n1 / 2 = row pairs
    int64_t ldi, ldo; // complex leading dimensions of in / out (half for the caller's matrix, padded for tmp)
    int NA, NB;
    const cplx* twL;  // roots of the FFT length of this kernel
    const cplx* wn;   // exp(-2 pi i m / n), m < n
    const cplx* wq;   // exp(-i pi k / (2n)), k <= n
    const cplx* wqB;  // exp(-i pi k0 / (2 NB)) = wq[NA k0], k0 <= NB: small, L1-resident
};

template <int NL, int CT_ = 128> struct RowsGeom {
    static constexpr int T = FftCfg<NL>::T;            // threads per line
    static constexpr int CT = CT_;                     // threads per CTA (>= 2 lines)
    static constexpr int LINES = CT / T;
    // consecutive lanes work on consecutive LINES (same element index): an odd slot stride spreads a quarter-warp's
    // 16-byte accesses over all eight bank groups
    static constexpr int STRIDE = NL + 1;
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)LINES * STRIDE;
};

// column of the real-pair matrix holding y[m]
__device__ __forceinline__ int64_t makhoul_col(int64_t m, int64_t n) { return (2 * m < n) ? 2 * m : 2 * (n - 1 - m) + 1; }

// ---- outer kernel: FFT_NA over m1 for fixed m0 = blockIdx.y ------------------------------------------------
//   forward: x[m1] = in[., col(m1 NB + m0)]            -> tmp[., k1 NB + m0] = w_n^(k1 m0) FFT(x)[k1]
//   inverse: x[k1] = tmp[., k1 NB + m0] (already times conj(w_n^(k1 m0)))  -> out[., col(m1 NB + m0)] = IFFT(x)[m1]
template <int NA, bool INV>
__global__ void __launch_bounds__(RowsGeom<NA>::CT) rows_outer_kernel(const __grid_constant__ RowsParams p) {
    typedef FftCfg<NA> C;
    typedef RowsGeom<NA> G;
    AFB_DYN_SMEM(cplx, smem);
    const int l = threadIdx.x % G::LINES, t = threadIdx.x / G::LINES;
    const int64_t ip = (int64_t)blockIdx.x * G::LINES + l;
    const bool act = ip < p.half;
    const int m0 = blockIdx.y;
    const int64_t n = (int64_t)NA * p.NB;
    cplx* s = smem + (size_t)l * G::STRIDE;
    // Loads are unconditional on a clamped row-pair index (inactive lines transform a copy of the last pair and
    // store nothing): a branch around them makes the compiler wait for each load before issuing the next.
    const cplx* src = p.in + (act ? ip : p.half - 1);
    cplx x[C::E];
    if (!INV) {
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = src[p.ldi * makhoul_col((int64_t)(t + C::T * r) * p.NB + m0, n)];
    } else {
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const cplx z = src[p.ldi * ((int64_t)(t + C::T * r) * p.NB + m0)];
            x[r] = make_double2(z.x, -z.y);  // inverse FFT = conj FFT conj (the twiddle was applied by the inner kernel)
        }
    }
    fft_line<NA>(s, t, x, p.twL);
    if (!act) return;
#pragma unroll
    for (int r = 0; r < C::E; ++r) {
        const int a = t + C::T * r;  // k1 (forward) / m1 (inverse)
        cplx z = line_get<NA>(s, a);
        if (!INV) {
            p.out[ip + p.ldo * ((int64_t)a * p.NB + m0)] = cmul(z, ldg2(p.wn + (int64_t)a * m0));
        } else {
            z.y = -z.y;
            p.out[ip + p.ldo * makhoul_col((int64_t)a * p.NB + m0, n)] = z;
        }
    }
}

// ---- inner kernel: groups k1 = g and NA - g (g = blockIdx.y <= NA/2) share a CTA, FFT_NB over m0 / k0 -------
//   forward: Y[k1 + NA k0] = FFT(tmp[., k1 NB + .])[k0], then the combine with Y[n-k] held by the partner line
//   inverse: Y built from c'[k], c'[n-k] (partner line), inverse FFT over k0, times conj(w_n^(k1 m0)) -> tmp[., k1 NB + m0]
template <int NB, bool INV>
__global__ void __launch_bounds__(INNER_CT) rows_inner_kernel(const __grid_constant__ RowsParams p) {
    typedef FftCfg<NB> C;
    typedef RowsGeom<NB, INNER_CT> G;
    constexpr int IB = G::LINES / 2;  // row pairs per CTA
    AFB_DYN_SMEM(cplx, smem);
    const int l = threadIdx.x % G::LINES, t = threadIdx.x / G::LINES;
    const int grp = l / IB, l2 = l % IB;
    const int64_t ip = (int64_t)blockIdx.x * IB + l2;
    const bool act = ip < p.half;
    const int g = blockIdx.y, NA = p.NA;
    const int k1 = grp ? (NA - g) % NA : g;
    const bool dup = grp && (g == 0 || 2 * g == NA);  // second copy of a self-paired group: computes, never stores
    const int64_t n = (int64_t)NA * NB;
    cplx* s = smem + (size_t)l * G::STRIDE;
    const cplx* sp = smem + (size_t)((1 - grp) * IB + l2) * G::STRIDE;  // partner line
    const cplx* src = p.in + (act ? ip : p.half - 1);  // see rows_outer_kernel
    cplx x[C::E];
    if (!INV) {
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = src[p.ldi * ((int64_t)k1 * NB + t + C::T * r)];
        fft_line<NB>(s, t, x, p.twL);
        if (!act || dup) return;
        // quarter-wave twiddle w^(k1 + NA k0) = w^k1 * (w^NA)^k0: one table entry per thread plus a 2 KB table that stays
        // in L1, instead of sixteen scattered reads of the (n+1)-entry table
        const cplx wk1 = ldg2(p.wq + k1);
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int k0 = t + C::T * r;
            const int64_t k = k1 + (int64_t)NA * k0;
            const int pk0 = (k1 == 0) ? (NB - k0) % NB : NB - 1 - k0;
            const cplx w = cmul(wk1, ldg2(p.wqB + k0));
            const cplx z = cadd(cmul(w, line_get<NB>(s, k0)), cmulc(line_get<NB>(sp, pk0), w));
            p.out[ip + p.ldo * k] = cscale(z, (k == 0) ? 0.5 / (double)n : 1.0 / (double)n);
        }
    } else {
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int k0 = t + C::T * r;
            x[r] = src[p.ldi * (k1 + (int64_t)NA * k0)];
            line_put<NB>(s, k0, x[r]);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int k0 = t + C::T * r;
            const int64_t k = k1 + (int64_t)NA * k0;
            const int pk0 = (k1 == 0) ? (NB - k0) % NB : NB - 1 - k0;
            cplx own = x[r], par = line_get<NB>(sp, pk0);
            if (k == 0) { own = cscale(own, 2.0); par = make_double2(0.0, 0.0); }
            const cplx d = make_double2(own.x + par.y, own.y - par.x);  // own - i par
            const cplx y = cscale(cmulc(d, ldg2(p.wq + k)), 0.5);  // (the factored twiddle of the forward kernel gave nothing here)
            x[r] = make_double2(y.x, -y.y);
        }
        __syncthreads();  // every partner read is done before the FFT reuses the buffer
        fft_line<NB>(s, t, x, p.twL);
        if (!act || dup) return;
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int m0 = t + C::T * r;
            // conj(result) * conj(w_n^(k1 m0)): the four-step twiddle is applied here, where no transform data is held in
            // registers (applying it in the outer kernel's load loop cost 64 more live registers and a resident CTA)
            const cplx z = cmul(line_get<NB>(s, m0), ldg2(p.wn + (int64_t)k1 * m0));
            p.out[ip + p.ldo * ((int64_t)k1 * NB + m0)] = make_double2(z.x, -z.y);
        }
    }
}

template <int NL, bool INV> static int32_t launch_outer(const RowsParams& p, void* stream) {
    typedef RowsGeom<NL> G;
    auto k = rows_outer_kernel<NL, INV>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    const int64_t gx = (p.half + G::LINES - 1) / G::LINES;
    AFB_LAUNCH(k, dim3((unsigned)gx, (unsigned)p.NB), G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
template <int NL, bool INV> static int32_t launch_inner(const RowsParams& p, void* stream) {
    typedef RowsGeom<NL, INNER_CT> G;
    auto k = rows_inner_kernel<NL, INV>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    const int64_t gx = (p.half + G::LINES / 2 - 1) / (G::LINES / 2);
    AFB_LAUNCH(k, dim3((unsigned)gx, (unsigned)(p.NA / 2 + 1)), G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
template <bool OUTER, bool INV> static int32_t launch_rows_n(int NL, const RowsParams& p, void* stream) {
#define AFB_CASE(N_) case N_: return OUTER ? launch_outer<N_, INV>(p, stream) : launch_inner<N_, INV>(p, stream)
    switch (NL) {
        AFB_CASE(32); AFB_CASE(64); AFB_CASE(128); AFB_CASE(256); AFB_CASE(512);
        default: return fail(AFB_UNSUPPORTED, "row transform: factor length not instantiated");
    }
#undef AFB_CASE
}

// n2 = NA * NB with NA >= NB, both powers of two in 32 .. 512
static bool rows_split(int64_t n2, int* NA, int* NB) {
    if (!is_pow2(n2)) return false;
    const int e = ilog2(n2);
    if (e < 10 || e > 18) return false;
    *NA = 1 << ((e + 1) / 2);
    *NB = 1 << (e / 2);
    return true;
}
bool rows_cheb_supported(int64_t n1, int64_t n2) {
    int NA, NB;
    return n1 >= 2 && n1 % 2 == 0 && n1 / 2 < ((int64_t)1 << 31) && rows_split(n2, &NA, &NB);
}

// The intermediate's columns are padded by 256 bytes: its column streams are read 128 at a time at a stride of NB
// columns, and an exact power-of-two byte stride between them was measured to halve the DRAM read rate.
int64_t rows_tmp_doubles(int64_t n1, int64_t n2) { return 2 * (n1 / 2 + ROWS_PAD) * n2; }

// in, out: n1 x n2 column-major (leading dimension n1), 16-byte aligned; in == out allowed;
// tmp: rows_tmp_doubles(n1, n2) doubles, distinct from both
int32_t launch_rows_cheb(bool inverse, const double* in, double* tmp, double* out, int64_t n1, int64_t n2, void* stream) {
    int NA = 0, NB = 0;
    if (!rows_cheb_supported(n1, n2) || !rows_split(n2, &NA, &NB))
        return fail(AFB_UNSUPPORTED, "row transform: unsupported shape");
    RowsParams p{};
    p.half = n1 / 2; p.NA = NA; p.NB = NB;
    const int64_t ldt = p.half + ROWS_PAD;
    const cplx *twA, *twB;
    AFB_TRY(root_table(NA, &twA));
    AFB_TRY(root_table(NB, &twB));
    AFB_TRY(root_table(n2, &p.wn));
    AFB_TRY(quarter_table(n2, &p.wq));
    AFB_TRY(quarter_table(NB, &p.wqB));
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* ctmp = reinterpret_cast<cplx*>(tmp);
    cplx* cout = reinterpret_cast<cplx*>(out);
    if (!inverse) {
        p.in = cin; p.ldi = p.half; p.out = ctmp; p.ldo = ldt; p.twL = twA;
        AFB_TRY((launch_rows_n<true, false>(NA, p, stream)));
        p.in = ctmp; p.ldi = ldt; p.out = cout; p.ldo = p.half; p.twL = twB;
        return launch_rows_n<false, false>(NB, p, stream);
    }
    p.in = cin; p.ldi = p.half; p.out = ctmp; p.ldo = ldt; p.twL = twB;
    AFB_TRY((launch_rows_n<false, true>(NB, p, stream)));
    p.in = ctmp; p.ldi = ldt; p.out = cout; p.ldo = p.half; p.twL = twA;
    return launch_rows_n<true, true>(NA, p, stream);
}

}  // namespace afb
