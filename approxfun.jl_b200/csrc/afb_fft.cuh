// Shared-memory Stockham FFT core (complex FP64, power-of-two length N = 32 .. 8192).
//
// One "line" = one length-N complex transform, worked on by T = N / R1 threads that each hold
// E = R1 complex values (16 for N >= 64) in registers.  A pass of radix R reads E values per thread
// from the line's shared-memory buffer (stride N/R apart: conflict free), multiplies by the pass
// twiddles, runs E/R register butterflies and scatters the results (autosort order) back.  The buffer
// is indexed through an XOR swizzle on 16-byte slots so that the first pass' scatter (stride R1 slots
// between lanes) is conflict free as well.
//
// Two drivers share the butterfly-level primitives:
//   DIT (forward kinds):  pass 1 from registers, ..., LAST pass leaves its outputs in registers.
//        The radix lists end in a radix <= E/2, so a thread runs >= 2 butterflies in the last pass;
//        they are assigned as partner pairs (j, NSL - j), which puts Z[k] and Z[N-k] in the SAME
//        thread: the real-FFT split / DCT quarter-wave post-processing then runs in registers and
//        goes straight to global memory (no shared-memory round trip for the post stage).
//   DIF (inverse kinds):  the exact transpose (passes in reverse order, butterfly then twiddle, load and
//        store patterns swapped), so the pre-processing of pairs (k, N-k) feeds the first pass from
//        registers and the last pass leaves z[col + T r] in registers.
//
// Twiddles: per butterfly one table load w = exp(-2 pi i k / (NS R)); the R-1 powers are formed in
// registers with a depth-<=4 product tree (the FP64 pipe has slack, the L1/shared datapath does not).
#pragma once
#include "afb_common.cuh"

namespace afb {

// ---- per-length configuration (DIT order) ---------------------------------------------------------
template <int N_, int R1_, int R2_, int R3_, int R4_> struct FftRadices {
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_, R4 = R4_;
    static constexpr int L = (R4_ > 1) ? 4 : (R3_ > 1) ? 3 : 2;
    static constexpr int T = N_ / R1_;
    static constexpr int E = R1_;
    static constexpr int SH = (R1_ == 32) ? 5 : (R1_ == 16) ? 4 : 3;   // log2 R1: the first pass' scatter stride
    static constexpr int RL = (L == 4) ? R4_ : (L == 3) ? R3_ : R2_;  // last radix
    static constexpr int NSL = N_ / RL;                               // butterflies in the last pass
    static constexpr int BL = E / RL;                                 // ... per thread
};
// Real kinds: the last radix is <= E/2, so every thread runs >= 2 butterflies of the last pass (partner pairs).
template <int N> struct FftCfg;
#define AFB_FFTCFG(N_, R1_, R2_, R3_, R4_) template <> struct FftCfg<N_> : FftRadices<N_, R1_, R2_, R3_, R4_> {}
AFB_FFTCFG(32, 8, 4, 1, 1);
AFB_FFTCFG(64, 16, 4, 1, 1);
AFB_FFTCFG(128, 16, 8, 1, 1);
AFB_FFTCFG(256, 16, 4, 4, 1);
AFB_FFTCFG(512, 16, 8, 4, 1);
AFB_FFTCFG(1024, 16, 8, 8, 1);
AFB_FFTCFG(2048, 16, 16, 8, 1);
#ifndef AFB_R32_4096
#define AFB_R32_4096 0
#endif
#if AFB_R32_4096
AFB_FFTCFG(4096, 32, 16, 8, 1);   // measurement variant: radix-32 first pass, 128 threads, three passes
#else
AFB_FFTCFG(4096, 16, 8, 4, 8);
#endif
// N = 8192 (real lines of 16384): ONE CTA owns an SM.  A radix-32 first pass (32 complex values per thread, 256 threads, three
// passes: 2 writes + 2 reads of the 128 KB exchange buffer instead of 3 + 3) was measured in round 2 (AFB_R32=1): of the HBM
// roofline at 16384 x 16384, 16.16.16.2 -> 32.16.16:  Chebyshev forward 0.634 -> 0.599, inverse 0.574 -> 0.643, Fourier
// forward 0.623 -> 0.663, inverse 0.609 -> 0.682.  A third less shared-memory traffic buys at most 12 %: the kernel is bound
// by the serialisation of its phases inside the single resident CTA, not by the exchange volume; the forward Chebyshev line
// (the column pass of the 2-D transform) loses, so the default stays 16.16.16.2.
#ifndef AFB_R32
#define AFB_R32 0
#endif
#if AFB_R32
AFB_FFTCFG(8192, 32, 16, 16, 1);
#else
AFB_FFTCFG(8192, 16, 16, 16, 2);
#endif
#undef AFB_FFTCFG
// Complex kinds (whole transform through shared memory, no pairing): fewest passes, largest radices first.  Same R1, so
// same T and E.  (The real kinds above prefer a LARGE last radix: fewer partner butterflies per thread means fewer twiddle
// and pair-table loads in the paired pass -- measured n = 1024: 0.83 -> 0.88, n = 2048 inverse 0.83 -> 0.89, n = 8192 inverse
// 0.60 -> 0.69 of HBM; for the complex kinds the same lists were slower, Laurent n = 4096 0.76 -> 0.66.)
template <int N> struct FftCfgC : FftCfg<N> {};
template <> struct FftCfgC<256> : FftRadices<256, 16, 16, 1, 1> {};
template <> struct FftCfgC<512> : FftRadices<512, 16, 16, 2, 1> {};
template <> struct FftCfgC<1024> : FftRadices<1024, 16, 16, 4, 1> {};
#if AFB_R32_4096
template <> struct FftCfgC<4096> : FftRadices<4096, 32, 16, 8, 1> {};
#else
template <> struct FftCfgC<4096> : FftRadices<4096, 16, 16, 8, 2> {};
#endif
// (4096 as 16.16.16 was measured slower than 16.16.8.2 for Laurent: 0.877 vs 0.847 ms at 32768 x 4096)

// Slots per line in shared memory.  Short lines are packed many per CTA and two of them share a quarter-warp: a stride
// of 4 (mod 8) slots keeps their 16-byte accesses on different bank groups (measured on one box, n = 64 / 128: +5..14 % over
// a stride of 0 mod 8; strides of 1, 2, 3, 6 mod 8 gave nothing).  N = 128 keeps its 8-slot pad, longer lines need none.
template <int N> struct FftLine {
    static constexpr int STRIDE = (N < 128) ? N + 4 : (N < 256) ? N + 8 : N;
};

template <int SH> __device__ __forceinline__ int swz(int i) { return i ^ ((i >> SH) & 7); }

// ---- register butterflies (forward sign, natural-order output) ------------------------------------
__device__ __forceinline__ void dft2(cplx& a, cplx& b) {
    const cplx t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
__device__ __forceinline__ void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
    const cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_mi(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

template <int R> struct Dft;
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(cplx* x) { dft2(x[0], x[1]); }
};
template <> struct Dft<4> {
    static __device__ __forceinline__ void run(cplx* x) { dft4(x[0], x[1], x[2], x[3]); }
};
template <> struct Dft<8> {
    static __device__ __forceinline__ void run(cplx* x) {
        const double h = 0.70710678118654752440;
        cplx e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
        cplx o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
        dft4(e0, e1, e2, e3);
        dft4(o0, o1, o2, o3);
        // w8^b * o_b
        const cplx p1 = make_double2((o1.x + o1.y) * h, (o1.y - o1.x) * h);
        const cplx p2 = cmul_mi(o2);
        const cplx p3 = make_double2((o3.y - o3.x) * h, -(o3.x + o3.y) * h);
        x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
        x[1] = cadd(e1, p1); x[5] = csub(e1, p1);
        x[2] = cadd(e2, p2); x[6] = csub(e2, p2);
        x[3] = cadd(e3, p3); x[7] = csub(e3, p3);
    }
};
template <> struct Dft<16> {
    static __device__ __forceinline__ void run(cplx* x) {
        const double h = 0.70710678118654752440;
        const double c1 = 0.92387953251128675613;  // cos(pi/8)
        const double s1 = 0.38268343236508977173;  // sin(pi/8)
        cplx y[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            y[a][0] = x[a]; y[a][1] = x[a + 4]; y[a][2] = x[a + 8]; y[a][3] = x[a + 12];
            dft4(y[a][0], y[a][1], y[a][2], y[a][3]);
        }
        // y[a][b] *= w16^(a b)
        y[1][1] = cmul(y[1][1], make_double2(c1, -s1));
        y[1][2] = make_double2((y[1][2].x + y[1][2].y) * h, (y[1][2].y - y[1][2].x) * h);
        y[1][3] = cmul(y[1][3], make_double2(s1, -c1));
        y[2][1] = make_double2((y[2][1].x + y[2][1].y) * h, (y[2][1].y - y[2][1].x) * h);
        y[2][2] = cmul_mi(y[2][2]);
        y[2][3] = make_double2((y[2][3].y - y[2][3].x) * h, -(y[2][3].x + y[2][3].y) * h);
        y[3][1] = cmul(y[3][1], make_double2(s1, -c1));
        y[3][2] = make_double2((y[3][2].y - y[3][2].x) * h, -(y[3][2].x + y[3][2].y) * h);
        y[3][3] = cmul(y[3][3], make_double2(-c1, s1));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            dft4(y[0][b], y[1][b], y[2][b], y[3][b]);
            x[b] = y[0][b]; x[b + 4] = y[1][b]; x[b + 8] = y[2][b]; x[b + 12] = y[3][b];
        }
    }
};

template <> struct Dft<32> {
    static __device__ __forceinline__ void run(cplx* x) {
        // w32^k = exp(-2 pi i k / 32), k = 1..15
        constexpr double C32[16] = {1.0, 0.98078528040323043058, 0.92387953251128673848, 0.83146961230254523567,
                                    0.70710678118654752440, 0.55557023301960228867, 0.38268343236508978178, 0.19509032201612833135,
                                    0.0, -0.19509032201612819257, -0.38268343236508972627, -0.55557023301960195560,
                                    -0.70710678118654752440, -0.83146961230254534669, -0.92387953251128673848, -0.98078528040323043058};
        constexpr double S32[16] = {0.0, 0.19509032201612824808, 0.38268343236508978178, 0.55557023301960217765,
                                    0.70710678118654752440, 0.83146961230254523567, 0.92387953251128673848, 0.98078528040323043058,
                                    1.0, 0.98078528040323043058, 0.92387953251128673848, 0.83146961230254545772,
                                    0.70710678118654752440, 0.55557023301960217765, 0.38268343236508989280, 0.19509032201612860891};
        cplx e[16], o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        Dft<16>::run(e);
        Dft<16>::run(o);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const cplx t = (k == 0) ? o[0] : (k == 8) ? cmul_mi(o[8]) : cmul(o[k], make_double2(C32[k], -S32[k]));
            x[k] = cadd(e[k], t);
            x[k + 16] = csub(e[k], t);
        }
    }
};

// multiply x[1..R-1] by w^1..w^(R-1)
template <int R> __device__ __forceinline__ void apply_powers(cplx* x, cplx w) {
    if constexpr (R == 2) {
        x[1] = cmul(x[1], w);
    } else {
        cplx p[R];
        p[1] = w;
#pragma unroll
        for (int r = 2; r < R; ++r) p[r] = cmul(p[r >> 1], p[r - (r >> 1)]);  // depth <= log2 R
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = cmul(x[r], p[r]);
    }
}

// ---- butterfly-level primitives.  Butterfly j of a pass (radix R, NS = product of earlier radices)
//      reads line elements   j + r N/R        (r < R)         "strided" pattern
//      writes line elements  (j-k) R + k + q NS, k = j mod NS  "autosort" pattern
template <int N, int R, int SH>
__device__ __forceinline__ void ld_strided(const cplx* __restrict__ s, int j, cplx* x) {
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = s[swz<SH>(j + r * (N / R))];
}
template <int N, int R, int SH>
__device__ __forceinline__ void st_strided(cplx* __restrict__ s, int j, const cplx* x) {
#pragma unroll
    for (int r = 0; r < R; ++r) s[swz<SH>(j + r * (N / R))] = x[r];
}
template <int N, int R, int NS, int SH>
__device__ __forceinline__ void ld_autosort(const cplx* __restrict__ s, int j, cplx* x) {
    const int k = j & (NS - 1);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = s[swz<SH>(base + q * NS)];
}
template <int N, int R, int NS, int SH>
__device__ __forceinline__ void st_autosort(cplx* __restrict__ s, int j, const cplx* x) {
    const int k = j & (NS - 1);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) s[swz<SH>(base + q * NS)] = x[q];
}
// DIT butterfly: twiddle then DFT.   DIF (transposed) butterfly: DFT then twiddle.
template <int N, int R, int NS>
__device__ __forceinline__ void bfly_dit(int j, cplx* x, const cplx* __restrict__ tw) {
    if constexpr (NS > 1) apply_powers<R>(x, ldg2(tw + (j & (NS - 1)) * (N / (NS * R))));
    Dft<R>::run(x);
}
template <int N, int R, int NS>
__device__ __forceinline__ void bfly_dif(int j, cplx* x, const cplx* __restrict__ tw) {
    Dft<R>::run(x);
    if constexpr (NS > 1) apply_powers<R>(x, ldg2(tw + (j & (NS - 1)) * (N / (NS * R))));
}

// ---- full passes with the default butterfly assignment j = t + b T ----------------------------------
template <int N, int R, int NS, class C = FftCfg<N>>
__device__ __forceinline__ void pass_dit(cplx* __restrict__ s, int t, cplx* x, const cplx* __restrict__ tw) {
    constexpr int B = C::E / R;
#pragma unroll
    for (int b = 0; b < B; ++b) ld_strided<N, R, C::SH>(s, t + b * C::T, x + b * R);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
        bfly_dit<N, R, NS>(t + b * C::T, x + b * R, tw);
        st_autosort<N, R, NS, C::SH>(s, t + b * C::T, x + b * R);
    }
    __syncthreads();
}
template <int N, int R, int NS>
__device__ __forceinline__ void pass_dif(cplx* __restrict__ s, int t, cplx* x, const cplx* __restrict__ tw) {
    typedef FftCfg<N> C;
    constexpr int B = C::E / R;
#pragma unroll
    for (int b = 0; b < B; ++b) ld_autosort<N, R, NS, C::SH>(s, t + b * C::T, x + b * R);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
        bfly_dif<N, R, NS>(t + b * C::T, x + b * R, tw);
        st_strided<N, R, C::SH>(s, t + b * C::T, x + b * R);
    }
    __syncthreads();
}

// ---- DIT driver pieces --------------------------------------------------------------------------------
// first pass: x[r] = input[col + T r] in registers  ->  shared memory (column `col` is free to choose)
template <int N, class C = FftCfg<N>>
__device__ __forceinline__ void dit_first(cplx* __restrict__ s, int col, cplx* x) {
    Dft<C::R1>::run(x);
    st_autosort<N, C::R1, 1, C::SH>(s, col, x);
    __syncthreads();
}
// passes 2 .. L-1
template <int N, class C = FftCfg<N>>
__device__ __forceinline__ void dit_middle(cplx* __restrict__ s, int t, cplx* x, const cplx* __restrict__ tw) {
    if constexpr (C::L >= 3) pass_dit<N, C::R2, C::R1, C>(s, t, x, tw);
    if constexpr (C::L >= 4) pass_dit<N, C::R3, C::R1 * C::R2, C>(s, t, x, tw);
}
// last pass on butterfly j: outputs x[q] = Z[j + NSL q] stay in registers
template <int N>
__device__ __forceinline__ void dit_last_bfly(const cplx* __restrict__ s, int j, cplx* x, const cplx* __restrict__ tw) {
    typedef FftCfg<N> C;
    ld_strided<N, C::RL, C::SH>(s, j, x);
    bfly_dit<N, C::RL, C::NSL>(j, x, tw);
}
// whole transform, result left in shared memory in natural order (complex kinds: Laurent, plain FFT, Bluestein, rows)
template <int N>
__device__ __forceinline__ void fft_line(cplx* __restrict__ s, int t, cplx* x, const cplx* __restrict__ tw) {
    typedef FftCfgC<N> C;
    dit_first<N, C>(s, t, x);
    dit_middle<N, C>(s, t, x, tw);
    pass_dit<N, C::RL, C::NSL, C>(s, t, x, tw);
}

// ---- DIF driver pieces (transpose of the above) -----------------------------------------------------------
// first executed = transposed last pass on butterfly j: x[q] = input[j + NSL q] from registers
template <int N>
__device__ __forceinline__ void dif_first_bfly(cplx* __restrict__ s, int j, cplx* x, const cplx* __restrict__ tw) {
    typedef FftCfg<N> C;
    bfly_dif<N, C::RL, C::NSL>(j, x, tw);
    st_strided<N, C::RL, C::SH>(s, j, x);
}
template <int N>
__device__ __forceinline__ void dif_middle(cplx* __restrict__ s, int t, cplx* x, const cplx* __restrict__ tw) {
    typedef FftCfg<N> C;
    if constexpr (C::L >= 4) pass_dif<N, C::R3, C::R1 * C::R2>(s, t, x, tw);
    if constexpr (C::L >= 3) pass_dif<N, C::R2, C::R1>(s, t, x, tw);
}
// final = transposed pass 1 on column col: outputs x[r] = result[col + T r] in registers
template <int N>
__device__ __forceinline__ void dif_last(const cplx* __restrict__ s, int col, cplx* x) {
    typedef FftCfg<N> C;
    ld_autosort<N, C::R1, 1, C::SH>(s, col, x);
    Dft<C::R1>::run(x);
}

// partner pairs of the last DIT / first DIF pass: pair index pi in [0, NSL/2):
//   pi >= 1: butterflies (pi, NSL - pi) ;  pi == 0: the two self-paired butterflies (0, NSL/2)
template <int N> __device__ __forceinline__ int pair_lo(int pi) { return pi; }
template <int N> __device__ __forceinline__ int pair_hi(int pi) {
    return pi == 0 ? FftCfg<N>::NSL / 2 : FftCfg<N>::NSL - pi;
}

// read / write element i of a natural-order line
template <int N> __device__ __forceinline__ cplx line_get(const cplx* s, int i) { return s[swz<FftCfg<N>::SH>(i)]; }
template <int N> __device__ __forceinline__ void line_put(cplx* s, int i, cplx v) { s[swz<FftCfg<N>::SH>(i)] = v; }

}  // namespace afb
