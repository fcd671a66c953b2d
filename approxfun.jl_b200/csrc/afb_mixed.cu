// Mixed-radix (2, 3, 4, 5, 7, 8) Stockham FFT in shared memory for lengths n = 2^a 3^b 5^c 7^d that are not powers of two:
// n = 1000, the plot sizes 3n + 50 of /root/reference/src/Plot/Plot.jl:16,24 when they are smooth, chopped coefficient
// counts, Padua grids.  One length-n complex FFT replaces the two padded (M >= 2n - 1) FFTs of Bluestein's algorithm.
//
//   * Radices are run-time data (any smooth n goes through ONE compiled kernel per kind); every pass is a switch onto a
//     compile-time butterfly.  The butterflies of all lines of a CTA are pooled and dealt to the threads round robin, so
//     odd butterfly counts (n / 5 = 200 on 256 threads) do not idle half a CTA.
//   * Autosort needs out-of-place passes when a thread runs a run-time number of butterflies: two buffers per line
//     (ping-pong), unpadded: stride-1 windows are conflict free at any alignment; see mx_pad.
//   * The four real kinds ride TWO real lines per complex FFT (z = line 2q + i line 2q+1), with the kind's pre-processing in
//     the load loop and the spectrum split + post-processing in the store loop, exactly as the fused Bluestein kernel
//     (afb_large.cu) does: one read and one write of the data, 16 B per coefficient.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "afb_fft.cuh"
#include "afb_internal.h"

namespace afb {

constexpr int MX_MAXPASS = 14;
constexpr int64_t MX_MAX_N = 6912;   // two buffers of 16 B * n: 216 KB at n = 6912

struct MixedSched {
    int n, npass, lpc, npad, ct, big;   // ct = threads per CTA (128 or 256); big = some pass has radix 16
    int u0c, u0f, ul, lpc_f;             // units per line pair: first pass (Chebyshev / Fourier pairing), last pass; lines per CTA of the forward kernel
    float inv_u0c, inv_u0f, inv_ul;
    int radix[MX_MAXPASS];
    int ns[MX_MAXPASS];          // product of the radices of the earlier passes
    float inv_nr[MX_MAXPASS];    // 1 / (n / radix)
    float inv_ns[MX_MAXPASS];
};

// Lines are NOT padded: every stride-1 window of eight 16-byte slots covers the 32 banks exactly once whatever its alignment
// (a pad slot inside an unaligned window made almost every access two wavefronts: ncu, round 2).  The only strided access
// is the scatter of the FIRST pass (stride = its radix), which is conflict free for an odd radix and 2-way for 6 and 10, so
// the schedule starts with such a radix.
__device__ __forceinline__ int mx_pad(int i) { return i; }
// floor(a / d) for 0 <= a < 2^20 through a float reciprocal, corrected to be exact
__device__ __forceinline__ int mx_div(int a, int d, float inv) {
    int q = (int)(((float)a + 0.5f) * inv);
    int r = a - q * d;
    if (r < 0) --q;
    else if (r >= d) ++q;
    return q;
}

// ---- odd-radix butterflies (forward sign): X_m = x0 + sum_k c_{mk} a_k - i sum_k s_{mk} b_k, a_k = x_k + x_{R-k}, b_k = x_k - x_{R-k}
// cos / sin of 2 pi i / R, i = 0..R-1 (compile-time index after unrolling, so these fold to immediates)
template <int R> __host__ __device__ constexpr double odd_cos(int i);
template <int R> __host__ __device__ constexpr double odd_sin(int i);
template <> __host__ __device__ constexpr double odd_cos<3>(int i) { return i == 0 ? 1.0 : -0.5; }
template <> __host__ __device__ constexpr double odd_sin<3>(int i) {
    return i == 0 ? 0.0 : i == 1 ? 0.86602540378443864676 : -0.86602540378443864676;
}
template <> __host__ __device__ constexpr double odd_cos<5>(int i) {
    return i == 0 ? 1.0 : (i == 1 || i == 4) ? 0.30901699437494742410 : -0.80901699437494742410;
}
template <> __host__ __device__ constexpr double odd_sin<5>(int i) {
    return i == 0 ? 0.0 : i == 1 ? 0.95105651629515357212 : i == 2 ? 0.58778525229247312917
         : i == 3 ? -0.58778525229247312917 : -0.95105651629515357212;
}
template <> __host__ __device__ constexpr double odd_cos<7>(int i) {
    return i == 0 ? 1.0 : (i == 1 || i == 6) ? 0.62348980185873353053 : (i == 2 || i == 5) ? -0.22252093395631440429
         : -0.90096886790241912624;
}
template <> __host__ __device__ constexpr double odd_sin<7>(int i) {
    return i == 0 ? 0.0 : i == 1 ? 0.78183148246802980871 : i == 2 ? 0.97492791218182360702 : i == 3 ? 0.43388373911755812048
         : i == 4 ? -0.43388373911755812048 : i == 5 ? -0.97492791218182360702 : -0.78183148246802980871;
}
template <int R> __device__ __forceinline__ void dft_odd(cplx* x) {
    constexpr int H = (R - 1) / 2;
    cplx a[H], b[H];
#pragma unroll
    for (int k = 0; k < H; ++k) { a[k] = cadd(x[k + 1], x[R - 1 - k]); b[k] = csub(x[k + 1], x[R - 1 - k]); }
    cplx x0 = x[0];
    cplx sum = x0;
#pragma unroll
    for (int k = 0; k < H; ++k) sum = cadd(sum, a[k]);
    x[0] = sum;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        cplx re = x0, im = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const double cc = odd_cos<R>(((m + 1) * (k + 1)) % R), ss = odd_sin<R>(((m + 1) * (k + 1)) % R);
            re.x = fma(cc, a[k].x, re.x); re.y = fma(cc, a[k].y, re.y);
            im.x = fma(ss, b[k].x, im.x); im.y = fma(ss, b[k].y, im.y);
        }
        // X_{m+1} = re - i im ; X_{R-1-m} = re + i im
        x[m + 1] = make_double2(re.x + im.y, re.y - im.x);
        x[R - 1 - m] = make_double2(re.x - im.y, re.y + im.x);
    }
}
// radix 2H (H = 3, 5): H-point transforms of the even and odd inputs, twiddle w_{2H}^k on the odd ones, radix-2 combine
template <int H> __device__ __forceinline__ void dft_two_odd(cplx* x) {
    cplx e[H], o[H];
#pragma unroll
    for (int i = 0; i < H; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
    dft_odd<H>(e);
    dft_odd<H>(o);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        // w_{2H}^k = -w_H^{(k + H) / 2} for odd k, w_H^{k/2} for even k
        const int idx = (k & 1) ? ((k + H) / 2) % H : k / 2;
        const double sg = (k & 1) ? -1.0 : 1.0;
        const cplx w = make_double2(sg * odd_cos<H>(idx), -sg * odd_sin<H>(idx));
        const cplx t = (k == 0) ? o[0] : cmul(o[k], w);
        x[k] = cadd(e[k], t);
        x[k + H] = csub(e[k], t);
    }
}
template <int R> struct MxDft { static __device__ __forceinline__ void run(cplx* x) { Dft<R>::run(x); } };
template <> struct MxDft<6> { static __device__ __forceinline__ void run(cplx* x) { dft_two_odd<3>(x); } };
template <> struct MxDft<10> { static __device__ __forceinline__ void run(cplx* x) { dft_two_odd<5>(x); } };
template <> struct MxDft<3> { static __device__ __forceinline__ void run(cplx* x) { dft_odd<3>(x); } };
template <> struct MxDft<5> { static __device__ __forceinline__ void run(cplx* x) { dft_odd<5>(x); } };
template <> struct MxDft<7> { static __device__ __forceinline__ void run(cplx* x) { dft_odd<7>(x); } };

// one pass of radix R over the pooled butterflies of the CTA's `lines` lines: src -> dst (both padded, stride `lstride` slots)
template <int R, int MX_CT>
__device__ __forceinline__ void mx_pass(const MixedSched& sc, int pass, const cplx* __restrict__ src, cplx* __restrict__ dst,
                                        int lines, int lstride, const cplx* __restrict__ tw) {
    const int n = sc.n, nr = n / R, NS = sc.ns[pass];
    const int twstep = n / (NS * R);
    const int total = lines * nr;
    for (int b = threadIdx.x; b < total; b += MX_CT) {
        const int l = (lines == 1) ? 0 : mx_div(b, nr, sc.inv_nr[pass]);
        const int j = b - l * nr;
        const cplx* s = src + l * lstride;
        cplx* d = dst + l * lstride;
        cplx x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = s[mx_pad(j + r * nr)];
        int k = 0;
        if (NS > 1) {
            k = j - mx_div(j, NS, sc.inv_ns[pass]) * NS;
            apply_powers<R>(x, ldg2(tw + k * twstep));
        }
        MxDft<R>::run(x);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) d[mx_pad(base + q * NS)] = x[q];
    }
}
// all passes; returns the buffer holding the natural-order result.  BIG = the schedule uses radix 16 (a separate instantiation:
// the radix-16 butterfly needs ~120 registers, the others fit 80 and keep three CTAs per SM)
template <bool BIG, int MX_CT>
__device__ __forceinline__ cplx* mx_fft(const MixedSched& sc, cplx* a, cplx* b, int lines, int lstride, const cplx* __restrict__ tw) {
    cplx *src = a, *dst = b;
    for (int p = 0; p < sc.npass; ++p) {
        switch (sc.radix[p]) {
            case 2: mx_pass<2, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 3: mx_pass<3, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 4: mx_pass<4, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 5: mx_pass<5, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 6: mx_pass<6, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 7: mx_pass<7, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 8: mx_pass<8, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            case 10: mx_pass<10, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
            default: if (BIG) mx_pass<16, MX_CT>(sc, p, src, dst, lines, lstride, tw); break;
        }
        __syncthreads();
        cplx* t = src; src = dst; dst = t;
    }
    return src;
}

// ---- complex lines: out[b*n + k] = scale * sum_j in[b*n + j] exp(-/+ 2 pi i jk/n) ----------------------------------
struct MixedCParams {
    const cplx* in;
    cplx* out;
    int64_t batch;
    const cplx* tw;
    int inverse;
    double scale;
    MixedSched sc;
};
template <bool BIG, int MX_CT>
__global__ void __launch_bounds__(MX_CT, (BIG ? 2 : 3) * (256 / MX_CT)) mixed_cfft_kernel(const __grid_constant__ MixedCParams p) {
    AFB_DYN_SMEM(cplx, smem);
    const MixedSched& sc = p.sc;
    const int n = sc.n, lstride = sc.npad;
    const int64_t line0 = (int64_t)blockIdx.x * sc.lpc;
    const int lines = (int)((p.batch - line0 < sc.lpc) ? (p.batch - line0) : sc.lpc);
    cplx* A = smem;
    cplx* B = smem + (size_t)sc.lpc * lstride;
    for (int l = 0; l < lines; ++l) {
        const cplx* src = p.in + (line0 + l) * n;
        for (int j = threadIdx.x; j < n; j += MX_CT) {
            cplx z = src[j];
            if (p.inverse) z.y = -z.y;
            A[l * lstride + mx_pad(j)] = z;
        }
    }
    __syncthreads();
    const cplx* R = mx_fft<BIG, MX_CT>(sc, A, B, lines, lstride, p.tw);
    for (int l = 0; l < lines; ++l) {
        cplx* dst = p.out + (line0 + l) * n;
        for (int j = threadIdx.x; j < n; j += MX_CT) {
            cplx z = cscale(R[l * lstride + mx_pad(j)], p.scale);
            if (p.inverse) z.y = -z.y;
            dst[j] = z;
        }
    }
}

// ---- the four real kinds, two real lines per complex FFT ------------------------------------------------------------
struct MixedRParams {
    const double* in;
    double* out;
    int64_t batch, istride, ostride;  // batch = real lines
    const cplx *tw, *qtab;
    MixedSched sc;
};
// pre-processing value of element j of one real line (same formulae as generic_pre_value in afb_large.cu)
template <int KIND>
__device__ __forceinline__ cplx mx_pre(const double* __restrict__ v, int n, int nh, int j, const cplx* __restrict__ qtab) {
    cplx z = make_double2(0.0, 0.0);
    if (KIND == AFB_CHEB1_FWD) {
        z.x = (j < nh) ? v[2 * j] : v[2 * (n - 1 - j) + 1];
    } else if (KIND == AFB_CHEB1_INV) {       // X'_k = conj(w^k) (c'_k - i c'_{n-k}) / 2,  c'_0 = 2 c_0, c'_n = 0
        if (j == 0) z.x = v[0];
        else {
            const cplx w = ldg2(qtab + j);
            z = cmul(make_double2(0.5 * v[j], -0.5 * v[n - j]), make_double2(w.x, -w.y));
        }
    } else if (KIND == AFB_FOURIER_FWD) {
        z.x = v[j];
    } else {                                  // AFB_FOURIER_INV
        if (j == 0) z.x = v[0];
        else if ((n % 2 == 0) && j == n / 2) z.x = -v[n - 1];
        else if (j < nh) z = make_double2(0.5 * v[2 * j], -0.5 * v[2 * j - 1]);
        else { const int k = n - j; z = make_double2(0.5 * v[2 * k], 0.5 * v[2 * k - 1]); }
    }
    return z;
}
struct MxLines {   // the two real lines of a pair
    const double2 *c0, *c1;
    double *o0, *o1;
    bool has2;
};
__device__ __forceinline__ MxLines mx_lines(const MixedRParams& p, int64_t q) {
    MxLines L;
    L.has2 = 2 * q + 1 < p.batch;
    const double* l0 = p.in + (2 * q) * p.istride;
    L.c0 = reinterpret_cast<const double2*>(l0);
    L.c1 = reinterpret_cast<const double2*>(L.has2 ? l0 + p.istride : l0);
    L.o0 = p.out + (2 * q) * p.ostride;
    L.o1 = L.o0 + p.ostride;
    return L;
}
// chunk c of both lines -> the complex values (even elements) and (odd elements)
__device__ __forceinline__ void mx_chunk(const MxLines& L, int c, cplx& ev, cplx& od) {
    const double2 a = L.c0[c];
    double2 b = L.c1[c];
    if (!L.has2) b = make_double2(0.0, 0.0);
    ev = make_double2(a.x, b.x);
    od = make_double2(a.y, b.y);
}
// post-processing of the pair A = W[k], B = W[n-k] (k == 0 or k == n/2: A == B)
template <int KIND>
__device__ __forceinline__ void mx_emit(const MxLines& L, const MixedRParams& p, int n, int k, cplx A, cplx B) {
    const double dn = (double)n;
    if (KIND == AFB_CHEB1_FWD) {
        const int kk[2] = {k, (k == 0) ? 0 : n - k};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (t == 1 && (kk[1] == kk[0] || k == 0)) break;
            const int o = kk[t];
            const cplx Av = t ? B : A, Bc = t ? A : B;
            const cplx Ye = make_double2(0.5 * (Av.x + Bc.x), 0.5 * (Av.y - Bc.y));    // (A + conj B) / 2
            const cplx Yo = make_double2(0.5 * (Av.y + Bc.y), -0.5 * (Av.x - Bc.x));   // (A - conj B) / (2i)
            const cplx w = ldg2(p.qtab + o);
            const double s = (o == 0) ? 1.0 / dn : 2.0 / dn;
            L.o0[o] = cmul(Ye, w).x * s;
            if (L.has2) L.o1[o] = cmul(Yo, w).x * s;
        }
    } else {   // AFB_FOURIER_FWD: frequency kf = min(k, n-k)
        const bool swap = (k != 0) && (2 * k > n);
        const int kf = swap ? n - k : k;
        const cplx Av = swap ? B : A, Bc = swap ? A : B;
        const cplx Ye = make_double2(0.5 * (Av.x + Bc.x), 0.5 * (Av.y - Bc.y));
        const cplx Yo = make_double2(0.5 * (Av.y + Bc.y), -0.5 * (Av.x - Bc.x));
        if (kf == 0) {
            L.o0[0] = Ye.x / dn;
            if (L.has2) L.o1[0] = Yo.x / dn;
        } else if (2 * kf == n) {
            L.o0[n - 1] = -Ye.x / dn;
            if (L.has2) L.o1[n - 1] = -Yo.x / dn;
        } else {
            L.o0[2 * kf - 1] = -Ye.y * (2.0 / dn);
            L.o0[2 * kf] = Ye.x * (2.0 / dn);
            if (L.has2) { L.o1[2 * kf - 1] = -Yo.y * (2.0 / dn); L.o1[2 * kf] = Yo.x * (2.0 / dn); }
        }
    }
}
template <int KIND, bool BIG, int MX_CT>
__global__ void __launch_bounds__(MX_CT, (BIG ? 2 : 3) * (256 / MX_CT)) mixed_real_kernel(const __grid_constant__ MixedRParams p) {
    constexpr bool INV = (KIND & 1) != 0;
    AFB_DYN_SMEM(cplx, smem);
    const MixedSched& sc = p.sc;
    const int n = sc.n, nh = (n + 1) / 2, lstride = sc.npad;
    const int64_t npairs = (p.batch + 1) / 2;
    const int64_t pair0 = (int64_t)blockIdx.x * sc.lpc;
    const int lines = (int)((npairs - pair0 < sc.lpc) ? (npairs - pair0) : sc.lpc);
    cplx* A = smem;
    cplx* B = smem + (size_t)sc.lpc * lstride;
    // forward kinds read every line as coalesced 128-bit chunks when the lines are 16-byte aligned: chunk m = (v[2m], v[2m+1])
    // feeds slots m and n-1-m (Chebyshev: Makhoul's even / odd split) or 2m and 2m+1 (Fourier)
    const bool vec = !INV && (n % 2 == 0) && (p.istride % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0);
    for (int l = 0; l < lines; ++l) {
        const int64_t q = pair0 + l;
        const bool has2 = 2 * q + 1 < p.batch;
        const double* l0 = p.in + (2 * q) * p.istride;
        const double* l1 = has2 ? l0 + p.istride : l0;
        cplx* Al = A + l * lstride;
        if (vec) {
            const double2* c0 = reinterpret_cast<const double2*>(l0);
            const double2* c1 = reinterpret_cast<const double2*>(l1);
            for (int m = threadIdx.x; m < n / 2; m += MX_CT) {
                const double2 a = c0[m];
                double2 b = c1[m];
                if (!has2) b = make_double2(0.0, 0.0);
                if (KIND == AFB_CHEB1_FWD) {
                    Al[mx_pad(m)] = make_double2(a.x, b.x);
                    Al[mx_pad(n - 1 - m)] = make_double2(a.y, b.y);
                } else {
                    Al[mx_pad(2 * m)] = make_double2(a.x, b.x);
                    Al[mx_pad(2 * m + 1)] = make_double2(a.y, b.y);
                }
            }
            continue;
        }
        if (INV) {
            // inverse kinds: elements j and n-j are built from the SAME coefficients (c_j, c_{n-j} / a_j, b_j), so one thread
            // makes both and every coefficient is loaded once (same arithmetic per element as mx_pre: same bits)
            for (int j = threadIdx.x; 2 * j <= n; j += MX_CT) {
                cplx za0, zb0, za1, zb1;   // line 0 / line 1, elements j (a) and n - j (b)
                const bool twin = j > 0 && 2 * j < n;
                if (KIND == AFB_CHEB1_INV) {
                    if (j == 0) { za0 = make_double2(l0[0], 0.0); za1 = make_double2(l1[0], 0.0); }
                    else {
                        const cplx wj = ldg2(p.qtab + j), wn = ldg2(p.qtab + (n - j));
                        const double c0j = l0[j], c0n = l0[n - j], c1j = l1[j], c1n = l1[n - j];
                        za0 = cmul(make_double2(0.5 * c0j, -0.5 * c0n), make_double2(wj.x, -wj.y));
                        za1 = cmul(make_double2(0.5 * c1j, -0.5 * c1n), make_double2(wj.x, -wj.y));
                        zb0 = cmul(make_double2(0.5 * c0n, -0.5 * c0j), make_double2(wn.x, -wn.y));
                        zb1 = cmul(make_double2(0.5 * c1n, -0.5 * c1j), make_double2(wn.x, -wn.y));
                    }
                } else {   // AFB_FOURIER_INV
                    if (j == 0) { za0 = make_double2(l0[0], 0.0); za1 = make_double2(l1[0], 0.0); }
                    else if (2 * j == n) { za0 = make_double2(-l0[n - 1], 0.0); za1 = make_double2(-l1[n - 1], 0.0); }
                    else {
                        const double a0 = l0[2 * j], b0 = l0[2 * j - 1], a1 = l1[2 * j], b1 = l1[2 * j - 1];
                        za0 = make_double2(0.5 * a0, -0.5 * b0); zb0 = make_double2(0.5 * a0, 0.5 * b0);
                        za1 = make_double2(0.5 * a1, -0.5 * b1); zb1 = make_double2(0.5 * a1, 0.5 * b1);
                    }
                }
                if (!has2) { za1 = make_double2(0.0, 0.0); zb1 = make_double2(0.0, 0.0); }
                Al[mx_pad(j)] = make_double2(za0.x - za1.y, -(za0.y + za1.x));            // conj(z0 + i z1)
                if (twin) Al[mx_pad(n - j)] = make_double2(zb0.x - zb1.y, -(zb0.y + zb1.x));
            }
            continue;
        }
        for (int j = threadIdx.x; j < n; j += MX_CT) {
            const cplx z1 = mx_pre<KIND>(l0, n, nh, j, p.qtab);
            cplx z2 = mx_pre<KIND>(l1, n, nh, j, p.qtab);
            if (!has2) z2 = make_double2(0.0, 0.0);
            cplx z = make_double2(z1.x - z2.y, z1.y + z2.x);  // z1 + i z2
            if (INV) z.y = -z.y;                              // inverse DFT = conj FFT conj
            Al[mx_pad(j)] = z;
        }
    }
    __syncthreads();
    const cplx* R = mx_fft<BIG, MX_CT>(sc, A, B, lines, lstride, p.tw);
    for (int l = 0; l < lines; ++l) {
        const int64_t q = pair0 + l;
        const cplx* W = R + l * lstride;
        const MxLines L = mx_lines(p, q);
        if (KIND == AFB_FOURIER_FWD) {
            // the pair W[k], W[n-k] gives the sine and cosine slots of frequency k of both lines: one thread reads each element
            // of the spectrum once (n = 1000: 0.521 -> 0.546 of HBM, n = 1536 0.508 -> 0.542)
            for (int k = threadIdx.x; 2 * k <= n; k += MX_CT) mx_emit<KIND>(L, p, n, k, W[mx_pad(k)], W[mx_pad(k == 0 ? 0 : n - k)]);
            continue;
        }
        if (KIND == AFB_CHEB1_FWD) {
            // one output per thread and step (the paired form, two outputs per pair, measured slower here: 0.532 -> 0.504)
            const double dn = (double)n;
            for (int o = threadIdx.x; o < n; o += MX_CT) {
                const cplx Av = W[mx_pad(o)], Bc = W[mx_pad(o == 0 ? 0 : n - o)];
                const cplx Ye = make_double2(0.5 * (Av.x + Bc.x), 0.5 * (Av.y - Bc.y));    // (A + conj B) / 2
                const cplx Yo = make_double2(0.5 * (Av.y + Bc.y), -0.5 * (Av.x - Bc.x));   // (A - conj B) / (2i)
                const cplx w = ldg2(p.qtab + o);
                const double sc2 = (o == 0) ? 1.0 / dn : 2.0 / dn;
                L.o0[o] = cmul(Ye, w).x * sc2;
                if (L.has2) L.o1[o] = cmul(Yo, w).x * sc2;
            }
            continue;
        }
        for (int o = threadIdx.x; o < n; o += MX_CT) {
            const cplx z = W[mx_pad(KIND == AFB_CHEB1_INV ? ((o & 1) ? n - 1 - (o - 1) / 2 : o / 2) : o)];
            L.o0[o] = z.x;
            if (L.has2) L.o1[o] = -z.y;
        }
    }
}

// =====================================================================================================================
// Forward kinds with the FIRST pass fed from global memory and the LAST pass finished in registers (round 2): the
// shared-memory traffic of a transform drops from 9 passes over the line (pre-store, 3 x load + store, two post loads at
// n = 1000) to 4 (first-pass store, one middle pass, last-pass load).
//   first pass: a thread owns the partner butterflies (j, nr-1-j) (Chebyshev: the 16-byte chunk (v[2m], v[2m+1]) holds
//               z[m] and its Makhoul mirror z[n-1-m], which are inputs r and R-1-r of the two partners) or (2a, 2a+1)
//               (Fourier: the chunk holds z[2c], z[2c+1]); every global access is a coalesced 128-bit load;
//   last pass:  a thread owns the partner butterflies (j, ns-j): their outputs are the pairs Z[k], Z[n-k] the spectrum
//               split and the kind's post-processing need, so both run in registers and go straight to global memory.
// Same arithmetic per output as mixed_real_kernel (bit-identical results).
// =====================================================================================================================
template <int R> __device__ __forceinline__ void mx_store_first(cplx* __restrict__ d, int j, cplx* x) {
    MxDft<R>::run(x);
#pragma unroll
    for (int q = 0; q < R; ++q) d[j * R + q] = x[q];
}
template <int KIND, int R, int MX_CT>
__device__ __forceinline__ void mx_first_fwd(const MixedSched& sc, const MixedRParams& p, int64_t pair0, int lines,
                                             cplx* __restrict__ dst, int lstride) {
    const int n = sc.n, nr = n / R;
    const int upl = (KIND == AFB_CHEB1_FWD) ? sc.u0c : sc.u0f;
    const float inv = (KIND == AFB_CHEB1_FWD) ? sc.inv_u0c : sc.inv_u0f;
    const int total = lines * upl;
    for (int b = threadIdx.x; b < total; b += MX_CT) {
        const int l = (lines == 1) ? 0 : mx_div(b, upl, inv);
        const int u = b - l * upl;
        const MxLines L = mx_lines(p, pair0 + l);
        cplx* d = dst + l * lstride;
        cplx xa[R], xb[R];
        if (KIND == AFB_CHEB1_FWD) {
            const int j = u, jp = nr - 1 - u;
            const bool self = jp == j;
            constexpr int H = R / 2;
#pragma unroll
            for (int r = 0; r < H; ++r) {
                mx_chunk(L, j + r * nr, xa[r], xb[R - 1 - r]);
                if (!self) mx_chunk(L, jp + r * nr, xb[r], xa[R - 1 - r]);
                else xa[R - 1 - r] = xb[R - 1 - r];
            }
            if (R & 1) mx_chunk(L, j + H * nr, xa[H], xb[H]);   // odd R: the middle inputs of both partners share a chunk
            mx_store_first<R>(d, j, xa);
            if (!self) mx_store_first<R>(d, jp, xb);
        } else {
            const int nr2 = nr / 2;
#pragma unroll
            for (int r = 0; r < R; ++r) mx_chunk(L, u + r * nr2, xa[r], xb[r]);
            mx_store_first<R>(d, 2 * u, xa);
            mx_store_first<R>(d, 2 * u + 1, xb);
        }
    }
}
template <int KIND, int R, int MX_CT>
__device__ __forceinline__ void mx_last_fwd(const MixedSched& sc, const MixedRParams& p, int64_t pair0, int lines,
                                            const cplx* __restrict__ src, int lstride) {
    const int n = sc.n, ns = n / R;
    const int upl = sc.ul;
    const int total = lines * upl;
    for (int b = threadIdx.x; b < total; b += MX_CT) {
        const int l = (lines == 1) ? 0 : mx_div(b, upl, sc.inv_ul);
        const int j = b - l * upl;
        if (2 * j > ns) continue;                 // ns odd: upl = (ns + 1) / 2 covers j = 0 .. (ns - 1) / 2 exactly; guard only
        const int jp = (j == 0) ? 0 : ns - j;
        const bool self = jp == j;
        const MxLines L = mx_lines(p, pair0 + l);
        const cplx* s = src + l * lstride;
        cplx xa[R], xb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xa[r] = s[j + r * ns];
        if (j > 0) apply_powers<R>(xa, ldg2(p.tw + j));
        MxDft<R>::run(xa);                        // xa[q] = Z[j + q ns]
        if (!self) {
#pragma unroll
            for (int r = 0; r < R; ++r) xb[r] = s[jp + r * ns];
            apply_powers<R>(xb, ldg2(p.tw + jp));
            MxDft<R>::run(xb);                    // xb[q] = Z[jp + q ns];  n - (j + q ns) = jp + (R-1-q) ns
#pragma unroll
            for (int q = 0; q < R; ++q) mx_emit<KIND>(L, p, n, j + q * ns, xa[q], xb[R - 1 - q]);
        } else if (j == 0) {                      // Z[q ns] pairs with Z[(R - q) ns]
            mx_emit<KIND>(L, p, n, 0, xa[0], xa[0]);
#pragma unroll
            for (int q = 1; 2 * q <= R; ++q) mx_emit<KIND>(L, p, n, q * ns, xa[q], xa[R - q]);
        } else {                                  // j = ns / 2: Z[j + q ns] pairs with Z[j + (R-1-q) ns]
#pragma unroll
            for (int q = 0; 2 * q <= R - 1; ++q) mx_emit<KIND>(L, p, n, j + q * ns, xa[q], xa[R - 1 - q]);
        }
    }
}
template <int KIND, bool BIG, int MX_CT>
// (two partner butterflies of radix 10 are 80 registers of data: 128 registers per thread, four 128-thread CTAs per SM)
__global__ void __launch_bounds__(MX_CT, (BIG ? 1 : 2) * (256 / MX_CT)) mixed_fwd_kernel(const __grid_constant__ MixedRParams p) {
    AFB_DYN_SMEM(cplx, smem);
    const MixedSched& sc = p.sc;
    const int lstride = sc.npad;
    const int64_t npairs = (p.batch + 1) / 2;
    const int64_t pair0 = (int64_t)blockIdx.x * sc.lpc_f;
    const int lines = (int)((npairs - pair0 < sc.lpc_f) ? (npairs - pair0) : sc.lpc_f);
    cplx* src = smem;
    cplx* dst = smem + (size_t)sc.lpc_f * lstride;
#define AFB_MX_FIRST(R_) case R_: mx_first_fwd<KIND, R_, MX_CT>(sc, p, pair0, lines, src, lstride); break
    switch (sc.radix[0]) {
        AFB_MX_FIRST(2); AFB_MX_FIRST(3); AFB_MX_FIRST(4); AFB_MX_FIRST(5); AFB_MX_FIRST(6); AFB_MX_FIRST(7); AFB_MX_FIRST(8);
        AFB_MX_FIRST(10);
        default: if (BIG) mx_first_fwd<KIND, 16, MX_CT>(sc, p, pair0, lines, src, lstride); break;
    }
#undef AFB_MX_FIRST
    __syncthreads();
    for (int ps = 1; ps + 1 < sc.npass; ++ps) {
        switch (sc.radix[ps]) {
            case 2: mx_pass<2, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 3: mx_pass<3, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 4: mx_pass<4, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 5: mx_pass<5, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 6: mx_pass<6, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 7: mx_pass<7, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 8: mx_pass<8, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            case 10: mx_pass<10, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
            default: if (BIG) mx_pass<16, MX_CT>(sc, ps, src, dst, lines, lstride, p.tw); break;
        }
        __syncthreads();
        cplx* t = src; src = dst; dst = t;
    }
#define AFB_MX_LAST(R_) case R_: mx_last_fwd<KIND, R_, MX_CT>(sc, p, pair0, lines, src, lstride); break
    switch (sc.radix[sc.npass - 1]) {
        AFB_MX_LAST(2); AFB_MX_LAST(3); AFB_MX_LAST(4); AFB_MX_LAST(5); AFB_MX_LAST(6); AFB_MX_LAST(7); AFB_MX_LAST(8);
        AFB_MX_LAST(10);
        default: if (BIG) mx_last_fwd<KIND, 16, MX_CT>(sc, p, pair0, lines, src, lstride); break;
    }
#undef AFB_MX_LAST
}

// ---- host side ----------------------------------------------------------------------------------------------------------
bool mixed_radix_ok(int64_t n) {
    if (n < 6 || n > MX_MAX_N || is_pow2(n)) return false;
    for (int f : {2, 3, 5, 7}) while (n % f == 0) n /= f;
    return n == 1;
}
// fewest passes over the available radices (depth-first over the divisors; n <= 6144 so this is instant)
static void mixed_factor(int64_t m, std::vector<int>& cur, std::vector<int>& best) {
    if (m == 1) { if (best.empty() || cur.size() < best.size()) best = cur; return; }
    if (!best.empty() && cur.size() + 1 >= best.size() && m > 16) return;
    for (int r : {16, 10, 8, 6, 7, 5, 4, 3, 2}) {
        if (m % r) continue;
        cur.push_back(r);
        mixed_factor(m / r, cur, best);
        cur.pop_back();
    }
}
static bool mixed_schedule(int64_t n, MixedSched* sc) {
    if (!mixed_radix_ok(n)) return false;
    std::vector<int> cur, r;
    mixed_factor(n, cur, r);
    if (r.empty() || (int)r.size() > MX_MAXPASS) return false;
    // largest radix first (the first pass has no twiddles) among those whose stride-R scatter is cheap: odd, else 6 / 10
    std::sort(r.begin(), r.end(), [](int a, int b) { return a > b; });
    auto rank_first = [](int R) { return (R & 1) ? 0 : (R == 6 || R == 10) ? 1 : 2; };
    size_t best = 0;
    for (size_t i = 1; i < r.size(); ++i)
        if (rank_first(r[i]) < rank_first(r[best])) best = i;
    std::swap(r[0], r[best]);
    std::sort(r.begin() + 1, r.end(), [](int a, int b) { return a > b; });
    sc->n = (int)n;
    sc->npass = (int)r.size();
    sc->npad = (int)n;
    int ns = 1;
    for (int p = 0; p < sc->npass; ++p) {
        sc->radix[p] = r[(size_t)p];
        sc->ns[p] = ns;
        sc->inv_nr[p] = 1.0f / (float)(n / r[(size_t)p]);
        sc->inv_ns[p] = 1.0f / (float)ns;
        ns *= r[(size_t)p];
    }
    // threads and lines per CTA: the pooled butterflies of the thinnest pass (n / largest radix per line) should fill ONE round
    // of the CTA's threads; short butterfly counts take 128-thread CTAs (twice as many resident CTAs in different phases).
    // Shared memory: ~36 KB per 128 threads (six / three CTAs per SM) when one line fits that at all; at most 16 lines.
    const int rmax = *std::max_element(r.begin(), r.end());
    sc->big = rmax == 16 ? 1 : 0;
    const int nr0 = (int)(n / rmax);
    sc->ct = (nr0 <= 160) ? 128 : 256;
    const size_t per_line = 2 * sizeof(cplx) * (size_t)sc->npad;
    const size_t budget = (sc->ct == 128) ? ((size_t)37 << 10) : ((size_t)74 << 10);
    const int fit = (int)std::max<size_t>(1, budget / per_line);
    const int want = std::max(1, sc->ct / nr0);
    sc->lpc = std::max(1, std::min({16, fit, want}));
    // register-resident first / last pass of the forward kinds: units = pairs of partner butterflies
    const int nr_first = (int)(n / r[0]), ns_last = (int)(n / r.back());
    sc->u0c = (nr_first + 1) / 2; sc->u0f = nr_first / 2; sc->ul = ns_last / 2 + 1;
    sc->inv_u0c = 1.0f / (float)sc->u0c; sc->inv_u0f = 1.0f / (float)std::max(sc->u0f, 1); sc->inv_ul = 1.0f / (float)sc->ul;
    // (packing more lines per CTA to fill the halved unit count of the paired passes was measured slower: two 96 KB CTAs per
    //  SM instead of six 32 KB ones; n = 1000 Fourier forward 0.417 -> 0.358 of HBM)
    sc->lpc_f = sc->lpc;
    return true;
}
// AFB_MIXED_SMEM_PASSES=1 keeps every pass in shared memory (the round-2 first version; A/B runs and tests)
static bool mixed_force_smem_passes() {
    static const bool v = [] { const char* e = std::getenv("AFB_MIXED_SMEM_PASSES"); return e && e[0] == '1'; }();
    return v;
}
static bool mixed_force_reg_passes() {
    static const bool v = [] { const char* e = std::getenv("AFB_MIXED_REG_PASSES"); return e && e[0] == '1'; }();
    return v;
}
static size_t mixed_smem(const MixedSched& sc) { return 2 * sizeof(cplx) * (size_t)sc.npad * (size_t)sc.lpc; }

int32_t launch_mixed_cfft(int64_t n, const cplx* tw, const cplx* in, cplx* out, int64_t batch, bool inverse, double scale,
                          void* stream) {
    if (batch == 0) return AFB_OK;
    MixedCParams p{in, out, batch, tw, inverse ? 1 : 0, scale, {}};
    if (!mixed_schedule(n, &p.sc)) return fail(AFB_UNSUPPORTED, "mixed-radix FFT: length is not 2^a 3^b 5^c 7^d <= 6912");
    const unsigned grid = (unsigned)((batch + p.sc.lpc - 1) / p.sc.lpc);
#define AFB_MX_LAUNCH(BIG_, CT_)                                               \
    do {                                                                       \
        auto k = mixed_cfft_kernel<BIG_, CT_>;                                 \
        AFB_SMEM_OPTIN(k, 220 * 1024);                                         \
        AFB_LAUNCH(k, grid, CT_, mixed_smem(p.sc), stream, p);                 \
    } while (0)
    if (p.sc.big) { if (p.sc.ct == 128) AFB_MX_LAUNCH(true, 128); else AFB_MX_LAUNCH(true, 256); }
    else { if (p.sc.ct == 128) AFB_MX_LAUNCH(false, 128); else AFB_MX_LAUNCH(false, 256); }
#undef AFB_MX_LAUNCH
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

template <int KIND> static int32_t launch_mixed_real_k(const MixedRParams& p, void* stream) {
    const int64_t npairs = (p.batch + 1) / 2;
    const unsigned grid = (unsigned)((npairs + p.sc.lpc - 1) / p.sc.lpc);
    // forward kinds: first pass from global memory, last pass in registers (needs 16-byte chunks that pair up: even n, even
    // line stride, aligned base, >= 2 passes; Fourier additionally an even butterfly count in the first pass)
    if constexpr (KIND == AFB_CHEB1_FWD || KIND == AFB_FOURIER_FWD) {
        const int n = p.sc.n;
        // Where it pays (measured on one box, fraction of the HBM roofline, shared-memory passes -> register passes):
        //   Fourier forward  n = 100 0.29 -> 0.43, 360 0.41 -> 0.50, 3000 0.37 -> 0.42, 5000 0.27 -> 0.33, 6000 0.26 -> 0.32,
        //                    but n = 1000 0.52 -> 0.42 and n = 1536 (radix 16: 230+ registers, one CTA per SM) 0.51 -> 0.28;
        //   Chebyshev forward loses everywhere (n = 1000 0.53 -> 0.25): two partner butterflies per thread halve the units of
        //   the first and last pass (50 on 128 threads) and cost 128 registers, so fewer, emptier CTAs hide less latency
        //   (ncu: warps active 10 %, issue 23 %, l1tex 19 %) -- the traffic saved was not the limiter.
        // AFB_MIXED_REG_PASSES=1 forces the register passes wherever the layout allows (tests, A/B runs).
        const bool layout = (n % 2 == 0) && (p.istride % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0) && p.sc.npass >= 2 &&
                            (KIND == AFB_CHEB1_FWD || (n / p.sc.radix[0]) % 2 == 0);
        const bool pays = KIND == AFB_FOURIER_FWD && !p.sc.big && (n <= 512 || n >= 2048);
        const bool ok = layout && !mixed_force_smem_passes() && (pays || mixed_force_reg_passes());
        if (ok) {
            const unsigned gridf = (unsigned)((npairs + p.sc.lpc_f - 1) / p.sc.lpc_f);
            const size_t smemf = 2 * sizeof(cplx) * (size_t)p.sc.npad * (size_t)p.sc.lpc_f;
#define AFB_MX_LAUNCH(BIG_, CT_)                                               \
    do {                                                                       \
        auto k = mixed_fwd_kernel<KIND, BIG_, CT_>;                            \
        AFB_SMEM_OPTIN(k, 220 * 1024);                                         \
        AFB_LAUNCH(k, gridf, CT_, smemf, stream, p);                           \
    } while (0)
            if (p.sc.big) { if (p.sc.ct == 128) AFB_MX_LAUNCH(true, 128); else AFB_MX_LAUNCH(true, 256); }
            else { if (p.sc.ct == 128) AFB_MX_LAUNCH(false, 128); else AFB_MX_LAUNCH(false, 256); }
#undef AFB_MX_LAUNCH
            AFB_KERNEL_CHECK();
            return AFB_OK;
        }
    }
#define AFB_MX_LAUNCH(BIG_, CT_)                                               \
    do {                                                                       \
        auto k = mixed_real_kernel<KIND, BIG_, CT_>;                           \
        AFB_SMEM_OPTIN(k, 220 * 1024);                                         \
        AFB_LAUNCH(k, grid, CT_, mixed_smem(p.sc), stream, p);                 \
    } while (0)
    if (p.sc.big) { if (p.sc.ct == 128) AFB_MX_LAUNCH(true, 128); else AFB_MX_LAUNCH(true, 256); }
    else { if (p.sc.ct == 128) AFB_MX_LAUNCH(false, 128); else AFB_MX_LAUNCH(false, 256); }
#undef AFB_MX_LAUNCH
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t launch_mixed_real(int kind, int64_t n, const cplx* tw, const cplx* qtab, const double* in, int64_t istride, double* out,
                          int64_t ostride, int64_t batch, void* stream) {
    if (batch == 0) return AFB_OK;
    MixedRParams p{in, out, batch, istride, ostride, tw, qtab, {}};
    if (!mixed_schedule(n, &p.sc)) return fail(AFB_UNSUPPORTED, "mixed-radix FFT: length is not 2^a 3^b 5^c 7^d <= 6912");
    switch (kind) {
        case AFB_CHEB1_FWD: return launch_mixed_real_k<AFB_CHEB1_FWD>(p, stream);
        case AFB_CHEB1_INV: return launch_mixed_real_k<AFB_CHEB1_INV>(p, stream);
        case AFB_FOURIER_FWD: return launch_mixed_real_k<AFB_FOURIER_FWD>(p, stream);
        case AFB_FOURIER_INV: return launch_mixed_real_k<AFB_FOURIER_INV>(p, stream);
        default: return fail(AFB_BAD_ARG, "mixed-radix real transform: kind");
    }
}

}  // namespace afb
