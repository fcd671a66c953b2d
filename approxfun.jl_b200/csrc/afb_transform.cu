// Batched 1-D transforms for power-of-two lengths that fit one CTA's shared memory
// (kernels K1, K2, K4, K5 of SURVEY.md section 2c) plus the O(n^2) direct kernels for tiny n.
//
//   CHEB1_FWD  (REDFT10,/n,c0/2):  c_0 = mean v, c_k = (2/n) sum_j v_j cos(pi k (j+1/2)/n)
//   CHEB1_INV  (REDFT01):          v_j = sum_k c_k cos(pi k (j+1/2)/n)
//   FOURIER_*  (R2HC/HC2R + interlace, /root/reference/src/Extras/fftGeneric.jl:47-64)
//   LAURENT_*  (c2c fft + [f0,f-1,f1,...] ordering, docs/src/usage/spaces.md:113-125)
//
// A length-n real transform runs as ONE complex FFT of length N = n/2 with everything else fused:
//   DCT-II : Makhoul's even/odd permutation is a lane-pair exchange on coalesced 128-bit loads feeding the
//            first pass from registers; the real-FFT split, the quarter-wave twiddle and the 1/n scaling run
//            in registers on the outputs of the last pass (partner butterflies (j, NSL-j) share a thread)
//            and go straight to global memory.  Two shared-memory exchanges per transform.
//   DCT-III: the transposed (DIF) flow: pair pre-processing in registers -> passes -> lane-pair exchange ->
//            coalesced 128-bit stores.       Fourier: same with the interlaced [a0,b1,a1,...] layout.
// Post/pre twiddles: one table load per pair-slot (index pi); the q-dependence k = pi + NSL q is a
// per-plan constant rotation passed in the kernel parameters (constant bank), so no per-element table traffic.
// Global traffic is exactly one read and one write of every element (16 B per coefficient).
#include <cstdlib>

#include "afb_fft.cuh"
#include "afb_internal.h"

namespace afb {

constexpr int MAXRL = LINE_MAXRL;   // largest last-pass radix
constexpr int MAXJ = 2;    // twiddle families per kind

struct LineParams {
    const void* in;
    void* out;
    int64_t batch;
    int64_t istride;  // elements between consecutive transforms (double; complex for Laurent / CFFT)
    int64_t ostride;
    const cplx* tw;   // exp(-2 pi i m / N), m < N
    const cplx* aux;  // real kinds: aux[J*pi + j] = t_j(pi), pi = 0..NSL/2, then t_j(N/2)
    double scale;
    // element strides (complex FFT both; Chebyshev kinds: output only, for the fused 2-D transpose)
    int64_t ies, oes;
    int conj_in, conj_out;
    int io_mode;      // LK_CFFT: bit 0 = input is a REAL array in Makhoul (DCT) order, bit 1 = output likewise
    int64_t bigN;     // ... of a length-bigN complex sequence z[m] = (v[4m], v[4m+2]) | (v[4q+3], v[4q+1]), q = bigN-1-m
    const cplx* twA;  // four-step: w_n^(m >> 12 << 12)
    const cplx* twB;  // w_n^(m & 4095);  output (line L, index i) is multiplied by w_n^((L mod twmod) * i)
    int64_t twmod;
    // batched four-step: line L = (b, c) with b = L / lpb; its first element sits at b * ibs + c * istride (lpb = 0: off)
    int64_t lpb, ibs, obs;
    // t_j(pi + NSL q) = t_j(pi) * K[j][q] (2q < RL);  t_j(N - pi - NSL q) = conj(t_j(pi)) * K[j][q] (2q >= RL)
    cplx K[MAXJ][MAXRL];
};

enum LineKind { LK_CHEB_FWD = 0, LK_CHEB_INV = 1, LK_FOURIER_FWD = 2, LK_FOURIER_INV = 3, LK_LAURENT_FWD = 4,
                LK_LAURENT_INV = 5, LK_CFFT = 6 };

template <int N> struct LineGeom {
    static constexpr int T = FftCfg<N>::T;
    static constexpr int CT = (T < 128) ? 128 : T;  // threads per CTA
    static constexpr int LPC = CT / T;              // lines per CTA
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)LPC * FftLine<N>::STRIDE;
};

// =================================================================================================
// Lane-pair exchange between the permuted DCT layout and first-pass columns.
//
// Makhoul's permutation packs a quad of reals v[4q..4q+3] into two complex FFT inputs
//     z[q] = (v[4q], v[4q+2])      z[N-1-q] = (v[4q+3], v[4q+1])            (q < N/2)
// Threads (2a, 2a+1) own first-pass columns (a, T-1-a).  For quad q the even lane loads the 16-byte
// chunk 2q and the odd lane chunk 2q+1 (fully coalesced 128-bit accesses); one 64-bit shuffle each way
// then gives z[q] to the owner of column q mod T and z[N-1-q] to its partner.
// =================================================================================================
template <int N> struct PairMap {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ int col(int t) { return (t & 1) ? C::T - 1 - (t >> 1) : (t >> 1); }
};

// HS > 0: the first HS of the R1/2 load groups (2T chunks each) come from the shared-memory stage st2.
template <int N, int HS = 0>
__device__ __forceinline__ void cheb_exchange_load(const double2* __restrict__ v2, bool act, int t, cplx* x,
                                                   const double2* st2 = nullptr) {
    typedef FftCfg<N> C;
    constexpr int H = C::R1 / 2;
    const int role = t & 1, p2 = t >> 1;
    double2 ch[H][2];
#pragma unroll
    for (int qq = 0; qq < H; ++qq)
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const int qcol = sg ? C::T - 1 - p2 : p2;
            const int q = qcol + C::T * qq;
            if (qq < HS) ch[qq][sg] = st2[2 * q + (role ^ sg)];
            else ch[qq][sg] = act ? v2[2 * q + (role ^ sg)] : make_double2(0.0, 0.0);
        }
#pragma unroll
    for (int qq = 0; qq < H; ++qq) {
        cplx z[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const int e = role ^ sg;
            const double send = e ? ch[qq][sg].x : ch[qq][sg].y;
            const double recv = __shfl_xor_sync(0xffffffffu, send, 1);
            z[sg] = make_double2(e ? ch[qq][sg].y : ch[qq][sg].x, recv);
        }
        x[qq] = role ? z[1] : z[0];
        x[C::R1 - 1 - qq] = role ? z[0] : z[1];
    }
}

// x[r] = r-th FFT output of column col(t); writes v = [Re z, Im z] un-permuted with z = conj(r)
template <int N>
__device__ __forceinline__ void cheb_exchange_store(double* __restrict__ v, int64_t oes, bool act, int t, const cplx* x) {
    typedef FftCfg<N> C;
    constexpr int H = C::R1 / 2;
    const int role = t & 1, p2 = t >> 1;
#pragma unroll
    for (int qq = 0; qq < H; ++qq) {
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const int qcol = sg ? C::T - 1 - p2 : p2;
            const int q = qcol + C::T * qq;
            const int e = role ^ sg;
            const cplx mine = e ? x[C::R1 - 1 - qq] : x[qq];
            const double recv = __shfl_xor_sync(0xffffffffu, mine.y, 1);
            // chunk 2q   (e = 0): (Re z[q], Im z[N-1-q]) = ( a.x, -b.y)
            // chunk 2q+1 (e = 1): (Im z[q], Re z[N-1-q]) = (-a.y,  b.x)
            const double2 out = e ? make_double2(-recv, mine.x) : make_double2(mine.x, -recv);
            if (act) {
                if (oes == 1) {
                    reinterpret_cast<double2*>(v)[2 * q + e] = out;
                } else {
                    const int64_t i0 = (int64_t)(4 * q + 2 * e);
                    v[i0 * oes] = out.x;
                    v[(i0 + 1) * oes] = out.y;
                }
            }
        }
    }
}

// =================================================================================================
// paired last pass (DIT) / first pass (DIF): visits every unordered pair {k, N-k} exactly once with
// both members in registers.  F (J twiddle families) provides
//     pair(k, A, B, t)  for 1 <= k < N/2 (A <-> index k, B <-> index N-k, t[j] = t_j(k)),
//     dc(Z0) and mid(Z_{N/2}, t) with t[j] = t_j(N/2).
// =================================================================================================
template <int J>
__device__ __forceinline__ void load_tw(const cplx* __restrict__ aux, int idx, cplx* tb) {
#pragma unroll
    for (int j = 0; j < J; ++j) tb[j] = ldg2(aux + J * idx + j);
}
// q is a compile-time constant after unrolling, so p.K[j][q] is a constant-bank operand
template <int N, int J>
__device__ __forceinline__ void rot_tw(const LineParams& p, const cplx* tb, int q, cplx* tq) {
    constexpr int RL = FftCfg<N>::RL;
#pragma unroll
    for (int j = 0; j < J; ++j) tq[j] = (2 * q < RL) ? cmul(tb[j], p.K[j][q]) : cmul(cconj(tb[j]), p.K[j][q]);
}

template <int N, int J, class F>
__device__ __forceinline__ void dit_last_paired(const LineParams& p, const cplx* __restrict__ s, int t, F& f) {
    typedef FftCfg<N> C;
    constexpr int RL = C::RL, NSL = C::NSL;
#pragma unroll
    for (int i = 0; i < C::BL / 2; ++i) {
        const int pi = t + C::T * i;
        cplx lo[RL], hi[RL], tb[J + 1], tq[J + 1];  // +1: J may be 0 (Laurent)
        dit_last_bfly<N>(s, pair_lo<N>(pi), lo, p.tw);
        dit_last_bfly<N>(s, pair_hi<N>(pi), hi, p.tw);
        load_tw<J>(p.aux, pi, tb);
        if (pi != 0) {
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                if (2 * q < RL) f.pair(pi + NSL * q, lo[q], hi[RL - 1 - q], tq);
                else f.pair((NSL - pi) + NSL * (RL - 1 - q), hi[RL - 1 - q], lo[q], tq);
            }
        } else {
            f.dc(lo[0]);
#pragma unroll
            for (int q = 1; q < RL / 2; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                f.pair(NSL * q, lo[q], lo[RL - q], tq);
            }
            load_tw<J>(p.aux, NSL / 2 + 1, tq);
            f.mid(lo[RL / 2], tq);
            load_tw<J>(p.aux, NSL / 2, tb);
#pragma unroll
            for (int q = 0; q < RL / 2; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                f.pair(NSL / 2 + NSL * q, hi[q], hi[RL - 1 - q], tq);
            }
        }
    }
}

// F additionally provides  fetch(k, raw) / fetch_dc(raw) (global loads only) so that every load of the
// thread is in flight before the first dependent FP64 instruction.
template <int N, int J, class F>
__device__ __forceinline__ void dif_first_paired(const LineParams& p, cplx* __restrict__ s, int t, F& f) {
    typedef FftCfg<N> C;
    constexpr int RL = C::RL, NSL = C::NSL, NP = C::BL / 2;
    typename F::Raw raw[NP][RL + 1];  // +1: the self-paired slot (pi == 0) holds dc, mid and RL-1 pairs
    // ---- phase A: issue every global load of the thread (twiddles are L1 hits and are fetched in phase B,
    //      prefetching them as well costs ~40 registers and was measured slower) -------------------------
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int pi = t + C::T * i;
        if (pi != 0) {
#pragma unroll
            for (int q = 0; q < RL; ++q)
                f.fetch((2 * q < RL) ? pi + NSL * q : (NSL - pi) + NSL * (RL - 1 - q), raw[i][q]);
        } else {
            f.fetch_dc(raw[i][0]);
#pragma unroll
            for (int q = 1; q < RL / 2; ++q) f.fetch(NSL * q, raw[i][q]);
            f.fetch(N / 2, raw[i][RL / 2]);
#pragma unroll
            for (int q = 0; q < RL / 2; ++q) f.fetch(NSL / 2 + NSL * q, raw[i][RL / 2 + 1 + q]);  // RL/2 slots: hi
        }
    }
    // ---- phase B: pair pre-processing + transposed last pass ------------------------------------------
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int pi = t + C::T * i;
        cplx lo[RL], hi[RL], tq[J + 1], tb[J + 1];
        load_tw<J>(p.aux, pi, tb);
        const cplx wlo = ldg2(p.tw + pair_lo<N>(pi));   // last-pass twiddle base w_N^j of both butterflies
        const cplx whi = ldg2(p.tw + pair_hi<N>(pi));
        if (pi != 0) {
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                if (2 * q < RL) f.pair(raw[i][q], lo[q], hi[RL - 1 - q], tq);
                else f.pair(raw[i][q], hi[RL - 1 - q], lo[q], tq);
            }
        } else {
            f.dc(raw[i][0], lo[0]);
#pragma unroll
            for (int q = 1; q < RL / 2; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                f.pair(raw[i][q], lo[q], lo[RL - q], tq);
            }
            cplx dummy;
            load_tw<J>(p.aux, NSL / 2 + 1, tq);
            f.pair(raw[i][RL / 2], lo[RL / 2], dummy, tq);
            load_tw<J>(p.aux, NSL / 2, tb);
#pragma unroll
            for (int q = 0; q < RL / 2; ++q) {
                rot_tw<N, J>(p, tb, q, tq);
                f.pair(raw[i][RL / 2 + 1 + q], hi[q], hi[RL - 1 - q], tq);
            }
        }
        // transposed last pass: DFT then twiddle by w_N^(j r), scatter to the strided positions
        Dft<RL>::run(lo);
        apply_powers<RL>(lo, wlo);
        st_strided<N, RL, C::SH>(s, pair_lo<N>(pi), lo);
        Dft<RL>::run(hi);
        apply_powers<RL>(hi, whi);
        st_strided<N, RL, C::SH>(s, pair_hi<N>(pi), hi);
    }
    __syncthreads();
}

template <int N, int KIND> struct LineOp;

// ---- Chebyshev forward (DCT-II) -------------------------------------------------------------------
// t_0(k) = w^k / n, t_1(k) = -i w^(5k) / n,  w = exp(-i pi / (2n)) ; scale = 1/n
template <int N> struct ChebFwdPost {
    double* c;
    int64_t es;
    double scale;
    bool act;
    __device__ __forceinline__ void pair(int k, cplx A, cplx B, const cplx* t) {
        constexpr int n = 2 * N;
        const cplx S = make_double2(A.x + B.x, A.y - B.y);   // A + conj B
        const cplx D = make_double2(A.x - B.x, A.y + B.y);   // A - conj B
        const cplx U = cmul(S, t[0]);
        const cplx V = cmul(D, t[1]);
        const cplx P = cadd(U, V), Q = csub(U, V);
        if (act) {
            c[(int64_t)k * es] = P.x;
            c[(int64_t)(n - k) * es] = -P.y;
            c[(int64_t)(N - k) * es] = (Q.x - Q.y) * 0.70710678118654752440;
            c[(int64_t)(N + k) * es] = (Q.x + Q.y) * 0.70710678118654752440;
        }
    }
    __device__ __forceinline__ void dc(cplx A) {
        if (act) {
            c[0] = (A.x + A.y) * scale;
            c[(int64_t)N * es] = (A.x - A.y) * (scale * 1.41421356237309504880);
        }
    }
    __device__ __forceinline__ void mid(cplx M, const cplx* t) {
        constexpr int n = 2 * N;
        const cplx S = make_double2(2.0 * M.x, 0.0);
        const cplx D = make_double2(0.0, 2.0 * M.y);
        const cplx P = cadd(cmul(S, t[0]), cmul(D, t[1]));
        if (act) {
            c[(int64_t)(N / 2) * es] = P.x;
            c[(int64_t)(n - N / 2) * es] = -P.y;
        }
    }
};
template <int N> struct LineOp<N, LK_CHEB_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_io(p, reinterpret_cast<const double*>(p.in) + (act ? line * p.istride : 0),
               reinterpret_cast<double*>(p.out) + (act ? line * p.ostride : 0), p.oes, act, t, s);
    }
    // the line at `in` -> the line at `out` (global memory, or the shared-memory tile of line_kernel_tile)
    static __device__ __forceinline__ void run_io(const LineParams& p, const double* in, double* out, int64_t oes, bool act, int t,
                                                  cplx* s) {
        const double2* v2 = reinterpret_cast<const double2*>(in);
        cplx x[C::E];
        cheb_exchange_load<N>(v2, act, t, x);
        dit_first<N>(s, PairMap<N>::col(t), x);
        dit_middle<N>(s, t, x, p.tw);
        ChebFwdPost<N> f{out, oes, p.scale, act};
        dit_last_paired<N, 2>(p, s, t, f);
    }
};

// ---- Chebyshev inverse (DCT-III) ------------------------------------------------------------------
// t_0(k) = conj(w^k), t_1(k) = exp(+2 pi i k / n)
template <int N> struct ChebInvPre {
    struct Raw { double ck, cnk, cNk, cNpk; };
    const double* c;
    bool act;
    __device__ __forceinline__ void fetch(int k, Raw& r) {
        constexpr int n = 2 * N;
        r.ck = 0; r.cnk = 0; r.cNk = 0; r.cNpk = 0;
        if (act) { r.ck = c[k]; r.cnk = c[n - k]; r.cNk = c[N - k]; r.cNpk = c[N + k]; }
    }
    __device__ __forceinline__ void fetch_dc(Raw& r) {
        r.ck = 0; r.cnk = 0; r.cNk = 0; r.cNpk = 0;
        if (act) { r.ck = c[0]; r.cNk = c[N]; }
    }
    __device__ __forceinline__ void pair(const Raw& r, cplx& ZK, cplx& ZNK, const cplx* t) {
        const double h = 0.70710678118654752440;
        const cplx wk = t[0];                                     // conj(w^k)
        const cplx G = cmul(make_double2(r.ck, -r.cnk), wk);     // X_k / N
        const cplx e = make_double2((wk.x + wk.y) * h, (wk.x - wk.y) * h);  // conj(w^(N-k))
        const cplx H = cmul(make_double2(r.cNk, -r.cNpk), e);    // X_{N-k} / N
        const cplx Xe = make_double2(0.5 * (G.x + H.x), 0.5 * (G.y - H.y));
        const cplx W = make_double2(0.5 * (G.x - H.x), 0.5 * (G.y + H.y));
        const cplx Xo = cmul(W, t[1]);
        ZK = make_double2(Xe.x - Xo.y, -Xe.y - Xo.x);            // conj(Z_k)     = conj(Xe) - i conj(Xo)
        ZNK = make_double2(Xe.x + Xo.y, Xe.y - Xo.x);            // conj(Z_{N-k}) = Xe - i Xo
    }
    __device__ __forceinline__ void dc(const Raw& r, cplx& Z0) {
        const double h = 0.70710678118654752440;
        Z0 = make_double2(r.ck + r.cNk * h, -(r.ck - r.cNk * h));
    }
};
// same with the first STAGED (> n/2) doubles of the line read from the shared-memory stage (k < N/2: c[k] and c[N-k] are
// always inside it, c[n-k] practically never, c[N+k] depends on STAGED)
template <int N, int STAGED> struct ChebInvPreStaged : ChebInvPre<N> {
    typedef typename ChebInvPre<N>::Raw Raw;
    const double* st;
    __device__ __forceinline__ ChebInvPreStaged(const double* c_, const double* st_) : ChebInvPre<N>{c_, true}, st(st_) {}
    __device__ __forceinline__ void fetch(int k, Raw& r) {
        constexpr int n = 2 * N;
        static_assert(STAGED > N, "stage must cover the first half of the line");
        r.ck = st[k]; r.cnk = this->c[n - k]; r.cNk = st[N - k];
        r.cNpk = (3 * N / 2 <= STAGED || N + k < STAGED) ? st[N + k] : this->c[N + k];
    }
    __device__ __forceinline__ void fetch_dc(Raw& r) {
        r.ck = st[0]; r.cnk = 0; r.cNk = st[N]; r.cNpk = 0;
    }
};
template <int N> struct LineOp<N, LK_CHEB_INV> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_io(p, reinterpret_cast<const double*>(p.in) + (act ? line * p.istride : 0),
               reinterpret_cast<double*>(p.out) + (act ? line * p.ostride : 0), p.oes, act, t, s);
    }
    static __device__ __forceinline__ void run_io(const LineParams& p, const double* in, double* out, int64_t oes, bool act, int t,
                                                  cplx* s) {
        ChebInvPre<N> f{in, act};
        dif_first_paired<N, 2>(p, s, t, f);
        cplx x[C::E];
        dif_middle<N>(s, t, x, p.tw);
        dif_last<N>(s, PairMap<N>::col(t), x);
        cheb_exchange_store<N>(out, oes, act, t, x);
    }
};

// ---- Fourier forward --------------------------------------------------------------------------------
// t_0(k) = -i exp(-2 pi i k / n) / n ;  scale = 1/n
template <int N> struct FourierFwdPost {
    double* c;
    double scale;
    bool act;
    __device__ __forceinline__ void pair(int k, cplx A, cplx B, const cplx* t) {
        const cplx S = make_double2((A.x + B.x) * scale, (A.y - B.y) * scale);
        const cplx D = make_double2(A.x - B.x, A.y + B.y);
        const cplx G = cmul(D, t[0]);
        if (act) {  // (2/n) F_k = S + G ; (2/n) F_{N-k} = conj(S - G)
            c[2 * k] = S.x + G.x;
            c[2 * k - 1] = -(S.y + G.y);
            c[2 * (N - k)] = S.x - G.x;
            c[2 * (N - k) - 1] = S.y - G.y;
        }
    }
    __device__ __forceinline__ void dc(cplx A) {
        if (act) {
            c[0] = (A.x + A.y) * scale;
            c[2 * N - 1] = -(A.x - A.y) * scale;
        }
    }
    __device__ __forceinline__ void mid(cplx M, const cplx* t) {
        const cplx S = make_double2(2.0 * M.x * scale, 0.0);
        const cplx D = make_double2(0.0, 2.0 * M.y);
        const cplx G = cmul(D, t[0]);
        if (act) {
            c[N] = S.x + G.x;
            c[N - 1] = -(S.y + G.y);
        }
    }
};
template <int N> struct LineOp<N, LK_FOURIER_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_io(p, reinterpret_cast<const double*>(p.in) + (act ? line * p.istride : 0),
               reinterpret_cast<double*>(p.out) + (act ? line * p.ostride : 0), 1, act, t, s);
    }
    static __device__ __forceinline__ void run_io(const LineParams& p, const double* in, double* out, int64_t, bool act, int t,
                                                  cplx* s) {
        const double2* v2 = reinterpret_cast<const double2*>(in);
        cplx x[C::E];
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = act ? v2[t + C::T * r] : make_double2(0.0, 0.0);
        dit_first<N>(s, t, x);
        dit_middle<N>(s, t, x, p.tw);
        FourierFwdPost<N> f{out, p.scale, act};
        dit_last_paired<N, 1>(p, s, t, f);
    }
};

// ---- Fourier inverse --------------------------------------------------------------------------------
// t_0(k) = exp(+2 pi i k / n)
template <int N> struct FourierInvPre {
    struct Raw { double a, b, aN, bN; };
    const double* c;
    bool act;
    __device__ __forceinline__ void fetch(int k, Raw& r) {
        r.a = 0; r.b = 0; r.aN = 0; r.bN = 0;
        if (act) { r.a = c[2 * k]; r.b = c[2 * k - 1]; r.aN = c[2 * (N - k)]; r.bN = c[2 * (N - k) - 1]; }
    }
    __device__ __forceinline__ void fetch_dc(Raw& r) {
        r.a = 0; r.b = 0; r.aN = 0; r.bN = 0;
        if (act) { r.a = c[0]; r.b = c[2 * N - 1]; }
    }
    __device__ __forceinline__ void pair(const Raw& r, cplx& ZK, cplx& ZNK, const cplx* t) {
        const cplx G = make_double2(r.a, -r.b);                   // F_k / N
        const cplx H = make_double2(r.aN, -r.bN);                 // F_{N-k} / N
        const cplx Xe = make_double2(0.5 * (G.x + H.x), 0.5 * (G.y - H.y));
        const cplx W = make_double2(0.5 * (G.x - H.x), 0.5 * (G.y + H.y));
        const cplx Xo = cmul(W, t[0]);
        ZK = make_double2(Xe.x - Xo.y, -Xe.y - Xo.x);
        ZNK = make_double2(Xe.x + Xo.y, Xe.y - Xo.x);
    }
    __device__ __forceinline__ void dc(const Raw& r, cplx& Z0) {
        Z0 = make_double2(r.a - r.b, -(r.a + r.b));  // F_0/N = 2 c0, F_N/N = -2 c_last
    }
};
// same with the first STAGED doubles of the line read from the shared-memory stage
template <int N, int STAGED> struct FourierInvPreStaged : FourierInvPre<N> {
    typedef typename FourierInvPre<N>::Raw Raw;
    const double* st;
    __device__ __forceinline__ FourierInvPreStaged(const double* c_, const double* st_) : FourierInvPre<N>{c_, true}, st(st_) {}
    __device__ __forceinline__ double ld(int i) const { return i < STAGED ? st[i] : this->c[i]; }
    __device__ __forceinline__ void fetch(int k, Raw& r) {
        r.a = st[2 * k]; r.b = st[2 * k - 1]; r.aN = ld(2 * (N - k)); r.bN = ld(2 * (N - k) - 1);  // k < N/2: 2k is staged
    }
    __device__ __forceinline__ void fetch_dc(Raw& r) {
        r.a = st[0]; r.b = this->c[2 * N - 1]; r.aN = 0; r.bN = 0;
    }
};
template <int N> struct LineOp<N, LK_FOURIER_INV> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_io(p, reinterpret_cast<const double*>(p.in) + (act ? line * p.istride : 0),
               reinterpret_cast<double*>(p.out) + (act ? line * p.ostride : 0), 1, act, t, s);
    }
    static __device__ __forceinline__ void run_io(const LineParams& p, const double* in, double* out, int64_t, bool act, int t,
                                                  cplx* s) {
        FourierInvPre<N> f{in, act};
        dif_first_paired<N, 1>(p, s, t, f);
        cplx x[C::E];
        dif_middle<N>(s, t, x, p.tw);
        dif_last<N>(s, t, x);
        if (!act) return;
        double2* v2 = reinterpret_cast<double2*>(out);
#pragma unroll
        for (int r = 0; r < C::E; ++r) v2[t + C::T * r] = make_double2(x[r].x, -x[r].y);
    }
};

// ---- Laurent / plain complex ------------------------------------------------------------------------
template <int N, class Op>
__device__ __forceinline__ void run_via_smem(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
    cplx x[FftCfg<N>::E];
    Op::load(p, line, act, t, x);
    fft_line<N>(s, t, x, p.tw);
    Op::store(p, line, act, t, s);
}
// (A register-resident variant through the paired last pass -- F_k and F_{N-k} stored as adjacent 16-byte slots -- was
// measured slower on the B200: 1.05 ms vs 0.84 ms at 32768 x 4096, its half-sector stores defeat coalescing.)
template <int N> struct LineOp<N, LK_LAURENT_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_via_smem<N, LineOp<N, LK_LAURENT_FWD>>(p, line, act, t, s);
    }
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* x) {
        const cplx* v = reinterpret_cast<const cplx*>(p.in) + (act ? line * p.istride : 0);
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = act ? v[t + C::T * r] : make_double2(0.0, 0.0);
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        cplx* c = reinterpret_cast<cplx*>(p.out) + line * p.ostride;
        for (int i = t; i < N; i += C::T) {  // (unrolling this loop was measured slower: 0.69 vs 0.79 of HBM)
            const int m = (i & 1) ? N - ((i + 1) >> 1) : (i >> 1);
            c[i] = cscale(line_get<N>(s, m), p.scale);  // scale = 1/n
        }
    }
};
template <int N> struct LineOp<N, LK_LAURENT_INV> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_via_smem<N, LineOp<N, LK_LAURENT_INV>>(p, line, act, t, s);
    }
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* x) {
        const cplx* c = reinterpret_cast<const cplx*>(p.in) + (act ? line * p.istride : 0);
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int m = t + C::T * r;
            const int i = (m < N / 2) ? 2 * m : 2 * (N - m) - 1;
            x[r] = act ? cconj(c[i]) : make_double2(0.0, 0.0);
        }
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        cplx* v = reinterpret_cast<cplx*>(p.out) + line * p.ostride;
        for (int j = t; j < N; j += C::T) v[j] = cconj(line_get<N>(s, j));
    }
};
// plain complex FFT with element strides (building block of Bluestein / four-step):
//   out[L*ostride + i*oes] = scale * tw(L,i) * FFT(in[L*istride + j*ies])_i,  optional conj on either side
template <int N> struct LineOp<N, LK_CFFT> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void run(const LineParams& p, int64_t line, bool act, int t, cplx* s) {
        run_via_smem<N, LineOp<N, LK_CFFT>>(p, line, act, t, s);
    }
    // Loads are unconditional (inactive tail lines re-read line 0 and store nothing) and the io_mode branch sits
    // outside the unrolled loop: a branch around each load makes the compiler wait for it before issuing the next.
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* x) {
        const int64_t l0 = act ? line : 0;
        if (p.io_mode & 1) {  // gather from the real array in DCT order
            const double* vr = reinterpret_cast<const double*>(p.in);
#pragma unroll
            for (int r = 0; r < C::E; ++r) {
                const int64_t m = l0 * p.istride + (int64_t)(t + C::T * r) * p.ies;
                const bool lo = 2 * m < p.bigN;
                const int64_t q = lo ? m : p.bigN - 1 - m;
                x[r] = make_double2(vr[4 * q + (lo ? 0 : 3)], vr[4 * q + (lo ? 2 : 1)]);
            }
        } else {
            int64_t off = l0 * p.istride;
            if (p.lpb) { const int64_t b = l0 / p.lpb; off = b * p.ibs + (l0 - b * p.lpb) * p.istride; }
            const cplx* v = reinterpret_cast<const cplx*>(p.in) + off;
#pragma unroll
            for (int r = 0; r < C::E; ++r) x[r] = v[(int64_t)(t + C::T * r) * p.ies];
        }
        if (p.conj_in) {
#pragma unroll
            for (int r = 0; r < C::E; ++r) x[r].y = -x[r].y;
        }
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        int64_t off = line * p.ostride;
        if (p.lpb) { const int64_t b = line / p.lpb; off = b * p.obs + (line - b * p.lpb) * p.ostride; }
        cplx* c = reinterpret_cast<cplx*>(p.out) + off;
        const int64_t lm = p.twA ? (line % p.twmod) : 0;
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int i = t + C::T * r;
            cplx z = line_get<N>(s, i);
            if (p.twA) {
                const int64_t m = lm * i;
                z = cmul(z, cmul(ldg2(p.twA + (m >> 12)), ldg2(p.twB + (m & 4095))));
            }
            z = cscale(z, p.scale);
            if (p.conj_out) z.y = -z.y;
            if (p.io_mode & 2) {  // scatter to the real array in DCT order
                double* vr = reinterpret_cast<double*>(p.out);
                const int64_t m = line * p.ostride + (int64_t)i * p.oes;
                if (2 * m < p.bigN) { vr[4 * m] = z.x; vr[4 * m + 2] = z.y; }
                else { const int64_t q = p.bigN - 1 - m; vr[4 * q + 3] = z.x; vr[4 * q + 1] = z.y; }
            } else {
                c[(int64_t)i * p.oes] = z;
            }
        }
    }
};

// =================================================================================================
// (Forcing 5 CTAs per SM on the forward kinds too -- 96 instead of 113 registers -- was measured slower: 0.887 vs
// 0.905 of HBM at n = 4096, 0.70 vs 0.81 at n = 256.)
template <int N, int KIND>
__global__ void __launch_bounds__(LineGeom<N>::CT, ((LineGeom<N>::CT == 128 && (KIND == LK_CHEB_INV || KIND == LK_FOURIER_INV)) ? (N == 4096 && AFB_R32_4096 ? 2 : 5) : (N == 8192 ? 1 : (N == 4096 && AFB_R32_4096 ? 2 : 512 / LineGeom<N>::CT)))) line_kernel(const __grid_constant__ LineParams p) {
    typedef LineGeom<N> G;
    AFB_DYN_SMEM(cplx, smem);
    const int tid = threadIdx.x;
    const int l = tid / G::T, t = tid % G::T;
    const int64_t line = (int64_t)blockIdx.x * G::LPC + l;
    const bool act = line < p.batch;
    cplx* s = smem + (size_t)l * FftLine<N>::STRIDE;
    LineOp<N, KIND>::run(p, line, act, t, s);
}

// -------------------------------------------------------------------------------------------------
// Short lines (n <= 256: four or eight threads per line, 16 or 32 lines per CTA).  The plain kernel's per-thread 8- and
// 16-byte accesses touch eight lines per warp instruction and keep the L1 84-99 % busy (ncu, profiles/
// r02_ncu_full_short_lines_summary.json).  Here the CTA's LPC contiguous lines move between global and shared memory as ONE
// tile with 128-bit accesses of consecutive threads, and the line code of LineOp runs on the tile (in place: a line's inputs
// are all in registers before the first barrier of its passes, its outputs are written after the last).
// Needs packed lines (istride = ostride = n, oes = 1).
// -------------------------------------------------------------------------------------------------
template <int N> struct TileGeom {
    typedef LineGeom<N> G;
    static constexpr int LINE = 2 * N;                                  // doubles per line
    static constexpr size_t FFT_BYTES = G::SMEM;
    static constexpr size_t SMEM = G::SMEM + sizeof(double) * (size_t)G::LPC * LINE;
};
template <int N, int KIND>
__global__ void __launch_bounds__(LineGeom<N>::CT, 3) line_kernel_tile(const __grid_constant__ LineParams p) {
    typedef LineGeom<N> G;
    typedef TileGeom<N> TG;
    AFB_DYN_SMEM(cplx, smem);
    double* tile = reinterpret_cast<double*>(smem + TG::FFT_BYTES / sizeof(cplx));
    const int tid = threadIdx.x;
    const int64_t line0 = (int64_t)blockIdx.x * G::LPC;
    const int nl = (int)((p.batch - line0 < G::LPC) ? (p.batch - line0) : G::LPC);      // valid lines of this CTA
    const int chunks = nl * (TG::LINE / 2);
    const double2* src = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(p.in) + line0 * TG::LINE);
    double2* t2 = reinterpret_cast<double2*>(tile);
    for (int i = tid; i < chunks; i += G::CT) t2[i] = src[i];
    __syncthreads();
    const int l = tid / G::T, t = tid % G::T;
    const bool act = l < nl;
    double* mine = tile + (size_t)(act ? l : 0) * TG::LINE;
    LineOp<N, KIND>::run_io(p, mine, mine, 1, act, t, smem + (size_t)l * FftLine<N>::STRIDE);
    __syncthreads();
    double2* dst = reinterpret_cast<double2*>(reinterpret_cast<double*>(p.out) + line0 * TG::LINE);
    for (int i = tid; i < chunks; i += G::CT) dst[i] = t2[i];
}

// -------------------------------------------------------------------------------------------------
// Long lines (N = 8192: one 512-thread CTA fills an SM's register file, so no second CTA can hide the load
// phase).  Persistent CTAs instead: while line i is transformed, the first 3/4 of the CTA's next line arrives
// in the spare 96 KB of shared memory through a TMA bulk copy (cp.async.bulk + mbarrier) and the last quarter is pulled into L2.
// -------------------------------------------------------------------------------------------------
template <int N> struct StageGeom {
    typedef FftCfg<N> C;
    // staged load groups (of R1/2 = 8) and resident CTAs per SM: N = 8192 stages 3/4 of the line beside its 128 KB
    // exchange buffer.  (Other N: whole line, three CTAs per SM -- measured at N = 2048: 0.77 of HBM against 0.905 for
    // the plain kernel with four CTAs per SM, so only N = 8192 is dispatched here.)
    static constexpr int HS = (N == 8192) ? 3 * (C::R1 / 2) / 4 : (N == 4096) ? 5 : 8;
    static constexpr int MINB = (N == 8192) ? 1 : (N == 4096) ? 2 : 3;
    static constexpr int FULL = 2 * N;
    // + the slot n*3/4 itself (k = N/2 of the inverse) when the stage is partial
    static constexpr int DOUBLES = (4 * C::T * HS + 16 < FULL) ? 4 * C::T * HS + 16 : FULL;
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)N + sizeof(double) * DOUBLES + 16;  // + the mbarrier
};
// One thread hands the 96 KB to the TMA engine (cp.async.bulk, completion counted on an mbarrier) and asks for the rest
// of the line in L2; no LSU instructions are spent on the copy.
template <int N>
__device__ __forceinline__ void stage_issue(double* __restrict__ st, const double* __restrict__ src, int t, uint64_t* bar) {
    typedef StageGeom<N> SG;
    if (t == 0) {
        bulk_copy_g2s(st, src, (uint32_t)(sizeof(double) * SG::DOUBLES), bar);
        if constexpr (SG::DOUBLES < SG::FULL)
            bulk_prefetch_l2(src + SG::DOUBLES, (uint32_t)(sizeof(double) * (SG::FULL - SG::DOUBLES)));
    }
}
template <int N, int KIND>
__global__ void __launch_bounds__(FftCfg<N>::T, StageGeom<N>::MINB) line_kernel_staged(const __grid_constant__ LineParams p) {
    typedef FftCfg<N> C;
    typedef StageGeom<N> SG;
    static_assert(KIND == LK_CHEB_FWD || KIND == LK_CHEB_INV || KIND == LK_FOURIER_FWD || KIND == LK_FOURIER_INV,
                  "staged lines: the real kinds");
    static_assert((sizeof(double) * SG::DOUBLES) % 16 == 0 && (sizeof(double) * (SG::FULL - SG::DOUBLES)) % 16 == 0, "bulk sizes");
    AFB_DYN_SMEM(cplx, s);
    double* st = reinterpret_cast<double*>(s + N);
    uint64_t& bar = *reinterpret_cast<uint64_t*>(st + SG::DOUBLES);
    const int t = threadIdx.x;
    const double* in = reinterpret_cast<const double*>(p.in);
    double* out = reinterpret_cast<double*>(p.out);
    if (t == 0) mbar_init(&bar);
    __syncthreads();
    uint32_t phase = 0;
    int64_t line = blockIdx.x;
    if (line < p.batch) stage_issue<N>(st, in + line * p.istride, t, &bar);
    for (; line < p.batch; line += gridDim.x) {
        const int64_t next = line + gridDim.x;
        mbar_wait(&bar, phase);  // the stage is complete (and visible: complete_tx)
        phase ^= 1;
        __syncthreads();         // every thread is done with the previous line's exchange buffer
        cplx x[C::E];
        if constexpr (KIND == LK_CHEB_FWD) {
            cheb_exchange_load<N, SG::HS>(reinterpret_cast<const double2*>(in + line * p.istride), true, t, x,
                                          reinterpret_cast<const double2*>(st));
            dit_first<N>(s, PairMap<N>::col(t), x);   // ends with a barrier: the stage is free again
            if (next < p.batch) stage_issue<N>(st, in + next * p.istride, t, &bar);
            dit_middle<N>(s, t, x, p.tw);
            ChebFwdPost<N> f{out + line * p.ostride, p.oes, p.scale, true};
            dit_last_paired<N, 2>(p, s, t, f);
        } else if constexpr (KIND == LK_CHEB_INV) {
            ChebInvPreStaged<N, SG::DOUBLES> f(in + line * p.istride, st);
            dif_first_paired<N, 2>(p, s, t, f);       // ends with a barrier
            if (next < p.batch) stage_issue<N>(st, in + next * p.istride, t, &bar);
            dif_middle<N>(s, t, x, p.tw);
            dif_last<N>(s, PairMap<N>::col(t), x);
            cheb_exchange_store<N>(out + line * p.ostride, p.oes, true, t, x);
        } else if constexpr (KIND == LK_FOURIER_FWD) {
            const double2* v2 = reinterpret_cast<const double2*>(in + line * p.istride);
            const double2* st2 = reinterpret_cast<const double2*>(st);
#pragma unroll
            for (int r = 0; r < C::E; ++r) x[r] = (2 * (C::T * (r + 1)) <= SG::DOUBLES) ? st2[t + C::T * r] : v2[t + C::T * r];
            dit_first<N>(s, t, x);
            if (next < p.batch) stage_issue<N>(st, in + next * p.istride, t, &bar);
            dit_middle<N>(s, t, x, p.tw);
            FourierFwdPost<N> f{out + line * p.ostride, p.scale, true};
            dit_last_paired<N, 1>(p, s, t, f);
        } else {
            FourierInvPreStaged<N, SG::DOUBLES> f(in + line * p.istride, st);
            dif_first_paired<N, 1>(p, s, t, f);
            if (next < p.batch) stage_issue<N>(st, in + next * p.istride, t, &bar);
            dif_middle<N>(s, t, x, p.tw);
            dif_last<N>(s, t, x);
            double2* v2 = reinterpret_cast<double2*>(out + line * p.ostride);
#pragma unroll
            for (int r = 0; r < C::E; ++r) v2[t + C::T * r] = make_double2(x[r].x, -x[r].y);
        }
    }
}
// -------------------------------------------------------------------------------------------------
// N = 8192 lines on a 2-CTA thread-block cluster (round 2; measured against line_kernel_staged).
// One line = 128 KB = half an SM's register file, so a single CTA per SM serialises its FP64 and shared-memory phases.
// Here CTA c of a cluster (c = 0, 1; two SMs) owns the samples z[2m + c] of a line: 256 threads, a 64 KB exchange buffer,
// three radix-16 passes of a 4096-point transform -- the first three passes of the 16.16.16.2 schedule never mix the two
// parity classes -- so two CTAs of DIFFERENT lines are resident per SM and overlap each other's phases.  The last (radix-2)
// pass pairs E[k] = FFT(even)[k] with O[k] = FFT(odd)[k]: it reads one operand from its own shared memory and the other
// from the partner's through distributed shared memory, and runs the same paired post-processing as dit_last_paired.
// Makhoul's permutation puts z[q] and z[N-1-q] (opposite parities) in one 32-byte sector, so each CTA uses half of every
// sector of the line (8-byte loads; the partner fetches the other half from L2 at about the same time).
// -------------------------------------------------------------------------------------------------
#ifndef AFB_HOST_EMU
}  // namespace afb
#include <cooperative_groups.h>
namespace afb {
namespace cg = cooperative_groups;

struct FftCfgHalf : FftRadices<4096, 16, 16, 16, 1> {};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// x[r] = z[2 (t + 256 r) + c]: r < 8 are the first half of z (even elements of a quad), r >= 8 the mirrored half
__device__ __forceinline__ void cluster_load_line(const double* __restrict__ v, int t, int c, cplx* x) {
    constexpr int N = 8192;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = 2 * (t + 256 * r) + c;
        x[r] = make_double2(v[4 * i], v[4 * i + 2]);
    }
#pragma unroll
    for (int r = 8; r < 16; ++r) {
        const int q = N - 1 - (2 * (t + 256 * r) + c);
        x[r] = make_double2(v[4 * q + 3], v[4 * q + 1]);
    }
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2) line_cluster_kernel(const __grid_constant__ LineParams p) {
    static_assert(KIND == LK_CHEB_FWD, "cluster lines: Chebyshev forward");
    constexpr int N = 8192, NH = 4096, NSL = 4096;
    typedef FftCfgHalf C;
    AFB_DYN_SMEM(cplx, s);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int t = threadIdx.x;
    const cplx* sE = cluster.map_shared_rank(s, 0);
    const cplx* sO = cluster.map_shared_rank(s, 1);
    const cplx* tw4 = p.twA;   // root table of 4096 (the sub-transforms); p.tw = root table of 8192 (the last pass)
    const int64_t nclusters = gridDim.x / 2;
    bool first = true;
    for (int64_t line = blockIdx.x / 2; line < p.batch; line += nclusters) {
        const double* v = reinterpret_cast<const double*>(p.in) + line * p.istride;
        // ---- first pass from registers; the global loads are in flight while the partner finishes reading this CTA's
        //      buffer (second half of the previous line's barrier pair)
        cplx x[C::E];
        cluster_load_line(v, t, c, x);
        if (!first) cluster_wait();
        first = false;
        dit_first<NH, C>(s, t, x);
        dit_middle<NH, C>(s, t, x, tw4);
        pass_dit<NH, C::RL, C::NSL, C>(s, t, x, tw4);     // Y_c in natural order
        cluster_arrive();                                  // this half is complete ...
        // ... meanwhile: pull the next line towards L2 and fetch this thread's post-processing twiddles
        if (line + nclusters < p.batch) {
            const double* vn = reinterpret_cast<const double*>(p.in) + (line + nclusters) * p.istride;
            if (t == 0) bulk_prefetch_l2(vn + 8192 * c, 65536u);   // each CTA asks for half of the 128 KB line
        }
        cplx tb[4][2];
#pragma unroll
        for (int g = 0; g < 4; ++g) load_tw<2>(p.aux, 1024 * c + t + 256 * g, tb[g]);
        cluster_wait();                                    // ... and so is the partner's
        // ---- last pass (radix 2 across the CTAs) + paired post-processing, pairs pi = 1024 c + t + 256 g
        ChebFwdPost<N> f{reinterpret_cast<double*>(p.out) + line * p.ostride, p.oes, p.scale, true};
        cplx lo[4][2], hi[4][2];
#pragma unroll
        for (int g = 0; g < 4; ++g) {                      // every shared / distributed-shared load first
            const int pi = 1024 * c + t + 256 * g;
            const int jh = (pi == 0) ? NSL / 2 : NSL - pi;
            lo[g][0] = sE[swz<C::SH>(pi)]; lo[g][1] = sO[swz<C::SH>(pi)];
            hi[g][0] = sE[swz<C::SH>(jh)]; hi[g][1] = sO[swz<C::SH>(jh)];
        }
        cluster_arrive();                                  // done reading the partner's buffer (waited for at the next line)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int pi = 1024 * c + t + 256 * g;
            const int jh = (pi == 0) ? NSL / 2 : NSL - pi;
            cplx tq[2];
            bfly_dit<N, 2, NSL>(pi, lo[g], p.tw);
            bfly_dit<N, 2, NSL>(jh, hi[g], p.tw);
            if (pi != 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) tq[j] = cmul(tb[g][j], p.K[j][0]);
                f.pair(pi, lo[g][0], hi[g][1], tq);
#pragma unroll
                for (int j = 0; j < 2; ++j) tq[j] = cmul(cconj(tb[g][j]), p.K[j][1]);
                f.pair(NSL - pi, hi[g][0], lo[g][1], tq);
            } else {
                f.dc(lo[g][0]);
                load_tw<2>(p.aux, NSL / 2 + 1, tq);
                f.mid(lo[g][1], tq);
                cplx tm[2];
                load_tw<2>(p.aux, NSL / 2, tm);
#pragma unroll
                for (int j = 0; j < 2; ++j) tq[j] = cmul(tm[j], p.K[j][0]);
                f.pair(NSL / 2, hi[g][0], hi[g][1], tq);
            }
        }
    }
    if (!first) cluster_wait();                            // no CTA exits while its partner may still read its shared memory
}
template <int KIND> static int32_t launch_line_cluster(const LineParams& p0, void* stream) {
    LineParams p = p0;
    AFB_TRY(root_table(4096, &p.twA));
    auto k = line_cluster_kernel<KIND>;
    const size_t smem = sizeof(cplx) * 4096;
    AFB_SMEM_OPTIN(k, smem);
    const int64_t cap = (int64_t)sm_count();              // two CTAs per SM = one cluster per SM on average
    const int64_t nclusters = p.batch < cap ? p.batch : cap;
    AFB_LAUNCH(k, (unsigned)(2 * nclusters), 256, smem, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
static bool line_cluster_enabled() {
    static const bool v = [] { const char* e = std::getenv("AFB_LINE_CLUSTER"); return e && e[0] == '1'; }();
    return v;
}
#endif

template <int N, int KIND> static int32_t launch_line_staged(const LineParams& p, void* stream) {
    typedef StageGeom<N> SG;
    auto k = line_kernel_staged<N, KIND>;
    AFB_SMEM_OPTIN(k, SG::SMEM);
    const int64_t cap = (int64_t)SG::MINB * sm_count();
    const int64_t grid = p.batch < cap ? p.batch : cap;
    AFB_LAUNCH(k, (unsigned)grid, FftCfg<N>::T, SG::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// AFB_LINE_TILE=1 / 0 switches the tile kernel for short lines on / off (default: see line_tile_enabled)
static bool line_tile_enabled() {
    static const int v = [] { const char* e = std::getenv("AFB_LINE_TILE"); return e ? (e[0] == '1' ? 1 : 0) : 0; }();
    return v != 0;
}

template <int N, int KIND> static int32_t launch_line_nk(const LineParams& p, void* stream) {
    typedef LineGeom<N> G;
    // (N = 4096 with two CTAs per SM and 5/8 of the line staged was measured SLOWER than the plain kernel: Chebyshev
    //  0.68 vs 0.74, Fourier 0.69 vs 0.75 of HBM -- two resident CTAs already overlap each other's load phase.)
    if constexpr (N == 8192 &&
                  (KIND == LK_CHEB_FWD || KIND == LK_CHEB_INV || KIND == LK_FOURIER_FWD || KIND == LK_FOURIER_INV)) {
        // 16-byte aligned contiguous lines and more lines than resident CTAs, otherwise there is nothing to overlap
#ifndef AFB_HOST_EMU
        if constexpr (KIND == LK_CHEB_FWD && !AFB_R32) {
            if (line_cluster_enabled() && p.batch >= 2) return launch_line_cluster<KIND>(p, stream);
        }
#endif
        if (p.oes == 1 && p.batch > (int64_t)StageGeom<N>::MINB * sm_count() && p.istride % 2 == 0 &&
            reinterpret_cast<uintptr_t>(p.in) % 16 == 0)
            return launch_line_staged<N, KIND>(p, stream);
    }
    if constexpr (N <= 128 && (KIND == LK_CHEB_FWD || KIND == LK_CHEB_INV || KIND == LK_FOURIER_FWD || KIND == LK_FOURIER_INV)) {
        // short packed lines: CTA-wide coalesced tile through shared memory (line_kernel_tile)
        if (line_tile_enabled() && p.oes == 1 && p.istride == 2 * N && p.ostride == 2 * N &&
            reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0) {
            auto kt = line_kernel_tile<N, KIND>;
            AFB_SMEM_OPTIN(kt, TileGeom<N>::SMEM);
            const int64_t gridt = (p.batch + G::LPC - 1) / G::LPC;
            AFB_LAUNCH(kt, (unsigned)gridt, G::CT, TileGeom<N>::SMEM, stream, p);
            AFB_KERNEL_CHECK();
            return AFB_OK;
        }
    }
    auto k = line_kernel<N, KIND>;
    AFB_SMEM_OPTIN(k, G::SMEM);
    const int64_t grid = (p.batch + G::LPC - 1) / G::LPC;
    AFB_LAUNCH(k, (unsigned)grid, G::CT, G::SMEM, stream, p);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

template <int KIND> static int32_t launch_line_k(int N, const LineParams& p, void* stream) {
    switch (N) {
        case 32: return launch_line_nk<32, KIND>(p, stream);
        case 64: return launch_line_nk<64, KIND>(p, stream);
        case 128: return launch_line_nk<128, KIND>(p, stream);
        case 256: return launch_line_nk<256, KIND>(p, stream);
        case 512: return launch_line_nk<512, KIND>(p, stream);
        case 1024: return launch_line_nk<1024, KIND>(p, stream);
        case 2048: return launch_line_nk<2048, KIND>(p, stream);
        case 4096: return launch_line_nk<4096, KIND>(p, stream);
        case 8192: return launch_line_nk<8192, KIND>(p, stream);
        default: return fail(AFB_UNSUPPORTED, "line FFT length not instantiated");
    }
}

// geometry of the last DIT pass for length N (host side: table / constant construction)
int32_t line_last_pass(int N, int* RL, int* NSL) {
#define AFB_CASE(N_) case N_: *RL = FftCfg<N_>::RL; *NSL = FftCfg<N_>::NSL; return AFB_OK
    switch (N) {
        AFB_CASE(32); AFB_CASE(64); AFB_CASE(128); AFB_CASE(256); AFB_CASE(512); AFB_CASE(1024); AFB_CASE(2048);
        AFB_CASE(4096); AFB_CASE(8192);
        default: return fail(AFB_UNSUPPORTED, "line FFT length not instantiated");
    }
#undef AFB_CASE
}

// N = complex FFT length; K = MAXJ x MAXRL rotation constants (may be null for the complex kinds)
int32_t launch_line(int kind, int N, const void* in, void* out, int64_t batch, int64_t istride, int64_t ostride,
                    int64_t oes, const cplx* tw, const cplx* aux, const cplx* K, double scale, void* stream) {
    if (batch == 0) return AFB_OK;
    LineParams p{};
    p.in = in; p.out = out; p.batch = batch; p.istride = istride; p.ostride = ostride; p.tw = tw; p.aux = aux;
    p.scale = scale; p.ies = 1; p.oes = oes; p.conj_in = 0; p.conj_out = 0; p.twA = nullptr; p.twB = nullptr; p.twmod = 1;
    if (K) for (int j = 0; j < MAXJ; ++j) for (int q = 0; q < MAXRL; ++q) p.K[j][q] = K[j * MAXRL + q];
    switch (kind) {
        case LK_CHEB_FWD: return launch_line_k<LK_CHEB_FWD>(N, p, stream);
        case LK_CHEB_INV: return launch_line_k<LK_CHEB_INV>(N, p, stream);
        case LK_FOURIER_FWD: return launch_line_k<LK_FOURIER_FWD>(N, p, stream);
        case LK_FOURIER_INV: return launch_line_k<LK_FOURIER_INV>(N, p, stream);
        case LK_LAURENT_FWD: return launch_line_k<LK_LAURENT_FWD>(N, p, stream);
        case LK_LAURENT_INV: return launch_line_k<LK_LAURENT_INV>(N, p, stream);
        default: return fail(AFB_BAD_ARG, "unknown line kind");
    }
}

// general strided complex FFT lines (see LineOp<N, LK_CFFT>)
int32_t launch_cfft_lines(int N, const cplx* in, cplx* out, int64_t nlines, int64_t istride, int64_t ies,
                          int64_t ostride, int64_t oes, const cplx* tw, bool conj_in, bool conj_out, double scale,
                          const cplx* twA, const cplx* twB, int64_t twmod, void* stream, int io_mode, int64_t bigN,
                          int64_t lpb, int64_t ibs, int64_t obs) {
    if (nlines == 0) return AFB_OK;
    LineParams p{};
    p.in = in; p.out = out; p.batch = nlines; p.istride = istride; p.ostride = ostride; p.tw = tw; p.aux = nullptr;
    p.scale = scale; p.ies = ies; p.oes = oes; p.conj_in = conj_in ? 1 : 0; p.conj_out = conj_out ? 1 : 0;
    p.twA = twA; p.twB = twB; p.twmod = twmod > 0 ? twmod : 1;
    p.io_mode = io_mode; p.bigN = bigN; p.lpb = lpb; p.ibs = ibs; p.obs = obs;
    return launch_line_k<LK_CFFT>(N, p, stream);
}

// =================================================================================================
// Direct O(n^2) kernels for tiny n (any n <= DIRECT_MAX): one thread per output element, exact integer
// argument reduction into a cosine / root table.
//   tab for Chebyshev : cos(pi m / (2n)), m = 0..4n-1
//   tab for Fourier/Laurent: exp(-2 pi i m / n), m = 0..n-1
// =================================================================================================
__global__ void direct_kernel(int kind, int n, const void* __restrict__ in, void* __restrict__ out, int64_t batch,
                              int64_t istride, int64_t ostride, int64_t oes, const cplx* __restrict__ tab) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * n) return;
    const int64_t line = gid / n;
    const int o = (int)(gid % n);
    const int l = n;
    if (kind == AFB_CHEB1_FWD || kind == AFB_CHEB1_INV) {
        const double* v = reinterpret_cast<const double*>(in) + line * istride;
        double* c = reinterpret_cast<double*>(out) + line * ostride;
        double acc = 0.0;
        if (kind == AFB_CHEB1_FWD) {
            for (int j = 0; j < n; ++j) acc += v[j] * tab[(o * (2 * j + 1)) % (4 * n)].x;
            c[(int64_t)o * oes] = acc * ((o == 0) ? 1.0 / n : 2.0 / n);
        } else {
            for (int k = 0; k < n; ++k) acc += v[k] * tab[(k * (2 * o + 1)) % (4 * n)].x;
            c[(int64_t)o * oes] = acc;
        }
    } else if (kind == AFB_FOURIER_FWD) {
        const double* v = reinterpret_cast<const double*>(in) + line * istride;
        double* c = reinterpret_cast<double*>(out) + line * ostride;
        // slot o: 0 -> a0 ; odd o=2k-1 -> b_k ; even o=2k -> a_k ; (even n) last -> -Re F_{n/2}/n
        int k;
        bool sine = false, nyq = false;
        if (o == 0) k = 0;
        else if ((l % 2 == 0) && o == l - 1) { k = l / 2; nyq = true; }
        else if (o & 1) { k = (o + 1) / 2; sine = true; }
        else k = o / 2;
        double re = 0.0, im = 0.0;
        for (int j = 0; j < n; ++j) {
            const cplx w = tab[(int)(((int64_t)j * k) % n)];
            re += v[j] * w.x;
            im += v[j] * w.y;
        }
        if (o == 0) c[o] = re / l;
        else if (nyq) c[o] = -re / l;
        else if (sine) c[o] = -im * (2.0 / l);
        else c[o] = re * (2.0 / l);
    } else if (kind == AFB_FOURIER_INV) {
        const double* c = reinterpret_cast<const double*>(in) + line * istride;
        double* v = reinterpret_cast<double*>(out) + line * ostride;
        // v_j = a0 + sum_k (a_k cos k th_j + b_k sin k th_j) - c_last cos(n/2 th_j)
        double acc = c[0];
        for (int i = 1; i < n; ++i) {
            if ((l % 2 == 0) && i == l - 1) {
                const cplx w = tab[(int)(((int64_t)o * (l / 2)) % n)];
                acc -= c[i] * w.x;
            } else {
                const int k = (i + 1) / 2;
                const cplx w = tab[(int)(((int64_t)o * k) % n)];
                acc += (i & 1) ? c[i] * (-w.y) : c[i] * w.x;  // tab = (cos, -sin)
            }
        }
        v[o] = acc;
    } else {  // Laurent
        const cplx* v = reinterpret_cast<const cplx*>(in) + line * istride;
        cplx* c = reinterpret_cast<cplx*>(out) + line * ostride;
        if (kind == AFB_LAURENT_FWD) {
            const int m = (o & 1) ? n - ((o + 1) >> 1) : (o >> 1);
            cplx acc = make_double2(0, 0);
            for (int j = 0; j < n; ++j) acc = cadd(acc, cmul(v[j], tab[(int)(((int64_t)j * m) % n)]));
            c[o] = cscale(acc, 1.0 / n);
        } else {
            cplx acc = make_double2(0, 0);
            for (int i = 0; i < n; ++i) {
                const int m = (i & 1) ? n - ((i + 1) >> 1) : (i >> 1);
                acc = cadd(acc, cmulc(v[i], tab[(int)(((int64_t)o * m) % n)]));
            }
            c[o] = acc;
        }
    }
}

int32_t launch_direct(int kind, int n, const void* in, void* out, int64_t batch, int64_t istride, int64_t ostride,
                      int64_t oes, const cplx* tab, void* stream) {
    if (batch == 0 || n == 0) return AFB_OK;
    const int64_t total = batch * n;
    auto k = direct_kernel;
    AFB_LAUNCH(k, (unsigned)((total + 127) / 128), 128, 0, stream, kind, n, in, out, batch, istride, ostride, oes, tab);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

}  // namespace afb
