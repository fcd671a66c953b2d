// Batched 1-D transforms for power-of-two lengths that fit one CTA's shared memory
// (kernels K1, K2, K4, K5 of SURVEY.md section 2c) plus the O(n^2) direct kernels for tiny n.
//
//   CHEB1_FWD  (REDFT10,/n,c0/2):  c_0 = mean v, c_k = (2/n) sum_j v_j cos(pi k (j+1/2)/n)
//   CHEB1_INV  (REDFT01):          v_j = sum_k c_k cos(pi k (j+1/2)/n)
//   FOURIER_*  (R2HC/HC2R + interlace, /root/reference/src/Extras/fftGeneric.jl:47-64)
//   LAURENT_*  (c2c fft + [f0,f-1,f1,...] ordering, docs/src/usage/spaces.md:113-125)
//
// A length-n real transform runs as ONE complex FFT of length N = n/2:
//   DCT-II : even/odd input permutation (Makhoul) folded into the first-pass register loads, real-FFT
//            split + quarter-wave twiddle + 1/n scaling folded into the store stage;
//   DCT-III: the exact mirror image;   Fourier: real-FFT split fused with the interlaced layout.
// Global traffic is exactly one read and one write of every element (16 B per coefficient).
#include "afb_fft.cuh"

namespace afb {

struct LineParams {
    const void* in;
    void* out;
    int64_t batch;
    int64_t istride;  // elements between consecutive transforms (double; complex for Laurent / CFFT)
    int64_t ostride;
    const cplx* tw;   // exp(-2 pi i m / N), m < N
    const cplx* aux;  // kind specific fused table
    double scale;
    // plain complex FFT (LK_CFFT) extras: element strides, conjugation flags, four-step twiddle
    int64_t ies, oes;
    int conj_in, conj_out;
    const cplx* twA;  // w_n^(m >> 12 << 12)
    const cplx* twB;  // w_n^(m & 4095);  output (line L, index i) is multiplied by w_n^((L mod twmod) * i)
    int64_t twmod;
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
// load stages: fill x[r] = z[t + T r] (first-pass inputs).  PRE kinds stage through shared memory.
// =================================================================================================
template <int N, int KIND> struct LineOp;

// ---- Chebyshev forward (DCT-II) -------------------------------------------------------------------
template <int N> struct LineOp<N, LK_CHEB_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        const double* v = reinterpret_cast<const double*>(p.in) + line * p.istride;
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            const int m = t + C::T * r;
            double re = 0.0, im = 0.0;
            if (act) {
                if (m < N / 2) {
                    re = v[4 * m];
                    im = v[4 * m + 2];
                } else {
                    const int q = N - 1 - m;
                    re = v[4 * q + 3];
                    im = v[4 * q + 1];
                }
            }
            x[r] = make_double2(re, im);
        }
    }
    // aux[2k] = w^k / n, aux[2k+1] = -i w^(5k) / n,  w = exp(-i pi / (2n)),  k = 0..N/2
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        constexpr int n = 2 * N;
        double* c = reinterpret_cast<double*>(p.out) + line * p.ostride;
        for (int k = 1 + t; k < N / 2; k += C::T) {
            const cplx A = line_get<N>(s, k), B = line_get<N>(s, N - k);
            const cplx S = make_double2(A.x + B.x, A.y - B.y);   // A + conj B
            const cplx D = make_double2(A.x - B.x, A.y + B.y);   // A - conj B
            const cplx U = cmul(S, ldg2(p.aux + 2 * k));
            const cplx V = cmul(D, ldg2(p.aux + 2 * k + 1));
            const cplx P = cadd(U, V), Q = csub(U, V);
            c[k] = P.x;
            c[n - k] = -P.y;
            c[N - k] = (Q.x - Q.y) * 0.70710678118654752440;
            c[N + k] = (Q.x + Q.y) * 0.70710678118654752440;
        }
        if (t == 0) {
            const cplx A = line_get<N>(s, 0);
            c[0] = (A.x + A.y) * p.scale;                             // scale = 1/n
            c[N] = (A.x - A.y) * (p.scale * 1.41421356237309504880);
            const cplx M = line_get<N>(s, N / 2);
            const cplx S = make_double2(2.0 * M.x, 0.0);
            const cplx D = make_double2(0.0, 2.0 * M.y);
            const cplx P = cadd(cmul(S, ldg2(p.aux + N)), cmul(D, ldg2(p.aux + N + 1)));
            c[N / 2] = P.x;
            c[n - N / 2] = -P.y;
        }
    }
};

// ---- Chebyshev inverse (DCT-III) ------------------------------------------------------------------
template <int N> struct LineOp<N, LK_CHEB_INV> {
    typedef FftCfg<N> C;
    // aux[2k] = conj(w^k), aux[2k+1] = exp(+2 pi i k / n),  k = 0..N/2
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        constexpr int n = 2 * N;
        const double* c = reinterpret_cast<const double*>(p.in) + line * p.istride;
        const double h = 0.70710678118654752440;
        for (int k = 1 + t; k <= N / 2; k += C::T) {
            double ck = 0, cnk = 0, cNk = 0, cNpk = 0;
            if (act) { ck = c[k]; cnk = c[n - k]; cNk = c[N - k]; cNpk = c[N + k]; }
            const cplx wk = ldg2(p.aux + 2 * k);                       // conj(w^k)
            const cplx G = cmul(make_double2(ck, -cnk), wk);           // X_k / N
            // conj(w^(N-k)) = e^{i pi/4} w^k = e^{i pi/4} conj(wk)
            const cplx e = make_double2((wk.x + wk.y) * h, (wk.x - wk.y) * h);
            const cplx H = cmul(make_double2(cNk, -cNpk), e);          // X_{N-k} / N
            const cplx Xe = make_double2(0.5 * (G.x + H.x), 0.5 * (G.y - H.y));
            const cplx W = make_double2(0.5 * (G.x - H.x), 0.5 * (G.y + H.y));
            const cplx Xo = cmul(W, ldg2(p.aux + 2 * k + 1));
            // conj(Z_k) = conj(Xe) - i conj(Xo) ;  conj(Z_{N-k}) = Xe - i Xo
            line_put<N>(s, k, make_double2(Xe.x - Xo.y, -Xe.y - Xo.x));
            line_put<N>(s, N - k, make_double2(Xe.x + Xo.y, Xe.y - Xo.x));
        }
        if (t == 0) {
            double c0 = 0, cN = 0;
            if (act) { c0 = c[0]; cN = c[N]; }
            const double a = c0 + cN * h, b = c0 - cN * h;
            line_put<N>(s, 0, make_double2(a, -b));
        }
        __syncthreads();
        pass_load<N, C::T, C::R1, C::SH>(s, t, x);
        __syncthreads();
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        double2* v2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(p.out) + line * p.ostride);
        for (int q = t; q < N / 2; q += C::T) {
            const cplx zq = line_get<N>(s, q), zr = line_get<N>(s, N - 1 - q);
            v2[2 * q] = make_double2(zq.x, -zr.y);
            v2[2 * q + 1] = make_double2(-zq.y, zr.x);
        }
    }
};

// ---- Fourier forward --------------------------------------------------------------------------------
template <int N> struct LineOp<N, LK_FOURIER_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        const double2* v2 = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(p.in) + line * p.istride);
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = act ? v2[t + C::T * r] : make_double2(0.0, 0.0);
    }
    // aux[k] = -i exp(-2 pi i k / n) / n,  k = 0..N/2 ;  scale = 1/n
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        constexpr int n = 2 * N;
        double* c = reinterpret_cast<double*>(p.out) + line * p.ostride;
        for (int k = 1 + t; k <= N / 2; k += C::T) {
            const cplx A = line_get<N>(s, k), B = line_get<N>(s, N - k);
            const cplx S = make_double2((A.x + B.x) * p.scale, (A.y - B.y) * p.scale);
            const cplx D = make_double2(A.x - B.x, A.y + B.y);
            const cplx G = cmul(D, ldg2(p.aux + k));
            // (2/n) F_k = S + G ; (2/n) F_{N-k} = conj(S - G)
            c[2 * k] = S.x + G.x;
            c[2 * k - 1] = -(S.y + G.y);
            if (k != N / 2) {
                c[2 * (N - k)] = S.x - G.x;
                c[2 * (N - k) - 1] = S.y - G.y;
            }
        }
        if (t == 0) {
            const cplx A = line_get<N>(s, 0);
            c[0] = (A.x + A.y) * p.scale;
            c[n - 1] = -(A.x - A.y) * p.scale;
        }
    }
};

// ---- Fourier inverse --------------------------------------------------------------------------------
template <int N> struct LineOp<N, LK_FOURIER_INV> {
    typedef FftCfg<N> C;
    // aux[k] = exp(+2 pi i k / n), k = 0..N/2
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        constexpr int n = 2 * N;
        const double* c = reinterpret_cast<const double*>(p.in) + line * p.istride;
        for (int k = 1 + t; k <= N / 2; k += C::T) {
            cplx G = make_double2(0, 0), H = make_double2(0, 0);
            if (act) {
                G = make_double2(c[2 * k], -c[2 * k - 1]);                // F_k / N
                H = make_double2(c[2 * (N - k)], -c[2 * (N - k) - 1]);    // F_{N-k} / N
            }
            const cplx Xe = make_double2(0.5 * (G.x + H.x), 0.5 * (G.y - H.y));
            const cplx W = make_double2(0.5 * (G.x - H.x), 0.5 * (G.y + H.y));
            const cplx Xo = cmul(W, ldg2(p.aux + k));
            line_put<N>(s, k, make_double2(Xe.x - Xo.y, -Xe.y - Xo.x));
            line_put<N>(s, N - k, make_double2(Xe.x + Xo.y, Xe.y - Xo.x));
        }
        if (t == 0) {
            double c0 = 0, cl = 0;
            if (act) { c0 = c[0]; cl = c[n - 1]; }
            // F_0/N = 2 c0, F_N/N = -2 c_last:  Z_0 = (c0 - cl) + i (c0 + cl)
            line_put<N>(s, 0, make_double2(c0 - cl, -(c0 + cl)));
        }
        __syncthreads();
        pass_load<N, C::T, C::R1, C::SH>(s, t, x);
        __syncthreads();
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        double2* v2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(p.out) + line * p.ostride);
        for (int m = t; m < N; m += C::T) {
            const cplx z = line_get<N>(s, m);
            v2[m] = make_double2(z.x, -z.y);
        }
    }
};

// ---- Laurent / plain complex ------------------------------------------------------------------------
template <int N> struct LineOp<N, LK_LAURENT_FWD> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        const cplx* v = reinterpret_cast<const cplx*>(p.in) + line * p.istride;
#pragma unroll
        for (int r = 0; r < C::E; ++r) x[r] = act ? v[t + C::T * r] : make_double2(0.0, 0.0);
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        cplx* c = reinterpret_cast<cplx*>(p.out) + line * p.ostride;
        for (int i = t; i < N; i += C::T) {
            const int m = (i & 1) ? N - ((i + 1) >> 1) : (i >> 1);
            c[i] = cscale(line_get<N>(s, m), p.scale);  // scale = 1/n
        }
    }
};
template <int N> struct LineOp<N, LK_LAURENT_INV> {
    typedef FftCfg<N> C;
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        const cplx* c = reinterpret_cast<const cplx*>(p.in) + line * p.istride;
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
    static __device__ __forceinline__ void load(const LineParams& p, int64_t line, bool act, int t, cplx* s, cplx* x) {
        const cplx* v = reinterpret_cast<const cplx*>(p.in) + line * p.istride;
#pragma unroll
        for (int r = 0; r < C::E; ++r) {
            cplx z = act ? v[(int64_t)(t + C::T * r) * p.ies] : make_double2(0.0, 0.0);
            if (p.conj_in) z.y = -z.y;
            x[r] = z;
        }
    }
    static __device__ __forceinline__ void store(const LineParams& p, int64_t line, bool act, int t, const cplx* s) {
        if (!act) return;
        cplx* c = reinterpret_cast<cplx*>(p.out) + line * p.ostride;
        const int64_t lm = p.twA ? (line % p.twmod) : 0;
        for (int i = t; i < N; i += C::T) {
            cplx z = line_get<N>(s, i);
            if (p.twA) {
                const int64_t m = lm * i;
                z = cmul(z, cmul(ldg2(p.twA + (m >> 12)), ldg2(p.twB + (m & 4095))));
            }
            z = cscale(z, p.scale);
            if (p.conj_out) z.y = -z.y;
            c[(int64_t)i * p.oes] = z;
        }
    }
};

// =================================================================================================
template <int N, int KIND>
__global__ void __launch_bounds__(LineGeom<N>::CT) line_kernel(LineParams p) {
    typedef LineGeom<N> G;
    AFB_DYN_SMEM(cplx, smem);
    const int tid = threadIdx.x;
    const int l = tid / G::T, t = tid % G::T;
    const int64_t line = (int64_t)blockIdx.x * G::LPC + l;
    const bool act = line < p.batch;
    cplx* s = smem + (size_t)l * FftLine<N>::STRIDE;
    cplx x[FftCfg<N>::E];
    LineOp<N, KIND>::load(p, line, act, t, s, x);
    fft_line<N>(s, t, x, p.tw);
    LineOp<N, KIND>::store(p, line, act, t, s);
}

template <int N, int KIND> static int32_t launch_line_nk(const LineParams& p, void* stream) {
    typedef LineGeom<N> G;
    auto k = line_kernel<N, KIND>;
    AFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
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

// N = complex FFT length
int32_t launch_line(int kind, int N, const void* in, void* out, int64_t batch, int64_t istride, int64_t ostride,
                    const cplx* tw, const cplx* aux, double scale, void* stream) {
    if (batch == 0) return AFB_OK;
    LineParams p{in, out, batch, istride, ostride, tw, aux, scale, 1, 1, 0, 0, nullptr, nullptr, 1};
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

// =================================================================================================
// Direct O(n^2) kernels for tiny n (any n <= DIRECT_MAX): one thread per output element, exact integer
// argument reduction into a cosine / root table.
//   tab for Chebyshev : cos(pi m / (2n)), m = 0..4n-1
//   tab for Fourier/Laurent: exp(-2 pi i m / n), m = 0..n-1
// =================================================================================================
__global__ void direct_kernel(int kind, int n, const void* __restrict__ in, void* __restrict__ out, int64_t batch,
                              int64_t istride, int64_t ostride, const cplx* __restrict__ tab) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * n) return;
    const int64_t line = gid / n;
    const int o = (int)(gid % n);
    const int l = n, nh = (n + 1) / 2;
    if (kind == AFB_CHEB1_FWD || kind == AFB_CHEB1_INV) {
        const double* v = reinterpret_cast<const double*>(in) + line * istride;
        double* c = reinterpret_cast<double*>(out) + line * ostride;
        double acc = 0.0;
        if (kind == AFB_CHEB1_FWD) {
            for (int j = 0; j < n; ++j) acc += v[j] * tab[(o * (2 * j + 1)) % (4 * n)].x;
            c[o] = acc * ((o == 0) ? 1.0 / n : 2.0 / n);
        } else {
            for (int k = 0; k < n; ++k) acc += v[k] * tab[(k * (2 * o + 1)) % (4 * n)].x;
            c[o] = acc;
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
        (void)nh;
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
                      const cplx* tab, void* stream) {
    if (batch == 0 || n == 0) return AFB_OK;
    const int64_t total = batch * n;
    auto k = direct_kernel;
    AFB_LAUNCH(k, (unsigned)((total + 127) / 128), 128, 0, stream, kind, n, in, out, batch, istride, ostride, tab);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

// general strided complex FFT lines (see LineOp<N, LK_CFFT>)
int32_t launch_cfft_lines(int N, const cplx* in, cplx* out, int64_t nlines, int64_t istride, int64_t ies,
                          int64_t ostride, int64_t oes, const cplx* tw, bool conj_in, bool conj_out, double scale,
                          const cplx* twA, const cplx* twB, int64_t twmod, void* stream) {
    if (nlines == 0) return AFB_OK;
    LineParams p{in, out, nlines, istride, ostride, tw, nullptr, scale, ies, oes, conj_in ? 1 : 0, conj_out ? 1 : 0,
                 twA, twB, twmod > 0 ? twmod : 1};
    return launch_line_k<LK_CFFT>(N, p, stream);
}

}  // namespace afb
