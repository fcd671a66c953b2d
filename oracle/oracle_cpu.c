/*
 * CPU oracle (plain C) for the Clenshaw / bisection / normalised-cumsum part of the ApproxFun
 * hot path.  TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product library.
 *
 * Every function restates a reference routine (paths are under /root/reference):
 *   - clenshaw recurrence: UPSTREAM ApproxFunOrthogonalPolynomials `clenshaw(::Chebyshev,c,x)`,
 *     in-tree call sites src/Extras/sample.jl:47,56,68,80,94.  Julia's `muladd` is restated as a
 *     true fma() (what LLVM emits on FMA hardware), so results are bit-comparable with the GPU's
 *     DFMA.  Parity status: pinned through tests/test_oracle_goldens.py (SamplingTest.jl:9,12,
 *     ReadmeTest.jl:23); the fused/unfused choice itself is unpinned in-tree.
 *   - bisection: src/Extras/sample.jl:41-52 (scalar), :58-75 (vector, k outer / i inner through
 *     the ClenshawPlan work vectors), :82-101 (matrix).
 *   - chebnormalizedcumsum!: src/Extras/sample.jl:121-175 + UPSTREAM ultraconversion!/ultraint!.
 *
 * Build: see oracle/Makefile (gcc -O2 -mfma -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- scalar Clenshaw (a7) ------------------------------------------------------------------- */
static inline double clenshaw1(const double* c, int64_t N, int64_t stride, double x) {
    if (N == 0) return 0.0;
    if (N == 1) return c[0];
    const double x2 = 2.0 * x;
    double b1 = 0.0, b2 = 0.0;
    for (int64_t k = N - 1; k >= 1; --k) {
        const double t = fma(x2, b1, c[k * stride] - b2);
        b2 = b1;
        b1 = t;
    }
    return fma(x2 / 2.0, b1, c[0] - b2);
}

/* point-outer evaluation: y[i] = sum_k c[k] T_k(x[i]) */
void orc_clenshaw(const double* c, int64_t N, const double* x, double* y, int64_t P, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < P; ++i) y[i] = clenshaw1(c, N, 1, x[i]);
}

/* Reference loop order (a8): k outer, i inner, three length-P work vectors owned by the plan.
 * UPSTREAM AFOP clenshaw(c::Vector, x::Vector, plan::ClenshawPlan); x is doubled in place and
 * restored, and the plan's buffer is what is returned.  work must hold 3*P doubles. */
void orc_clenshaw_planned(const double* c, int64_t N, double* x, double* y, int64_t P, double* work,
                          int nthreads) {
    double* bk = work;
    double* bk1 = work + P;
    double* bk2 = work + 2 * P;
    if (N == 0) { memset(y, 0, sizeof(double) * (size_t)P); return; }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
#pragma omp for schedule(static)
        for (int64_t i = 0; i < P; ++i) { x[i] = 2.0 * x[i]; bk1[i] = 0.0; bk2[i] = 0.0; }
        for (int64_t k = N - 1; k >= 1; --k) {
            const double ck = c[k];
#pragma omp for schedule(static)
            for (int64_t i = 0; i < P; ++i) {
                bk[i] = fma(x[i], bk1[i], ck - bk2[i]);
            }
#pragma omp for schedule(static)
            for (int64_t i = 0; i < P; ++i) { bk2[i] = bk1[i]; bk1[i] = bk[i]; }
        }
#pragma omp for schedule(static)
        for (int64_t i = 0; i < P; ++i) {
            x[i] = x[i] / 2.0;
            y[i] = (N == 1) ? c[0] : fma(x[i], bk1[i], c[0] - bk2[i]);
        }
    }
}

/* a9: column j of C (N x P, column-major, leading dimension ld) at x[j] */
void orc_clenshaw_cols(const double* C, int64_t N, int64_t ld, const double* x, double* y, int64_t P,
                       int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t j = 0; j < P; ++j) y[j] = clenshaw1(C + j * ld, N, 1, x[j]);
}

/* every column of C (N x M) at every x: Y is M x P row-major (Y[m*P + i]) */
void orc_clenshaw_multi(const double* C, int64_t N, int64_t M, const double* x, double* Y, int64_t P,
                        int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < P; ++i)
        for (int64_t m = 0; m < M; ++m) Y[m * P + i] = clenshaw1(C + m * N, N, 1, x[i]);
}

/* ---- general three-term Clenshaw (8f rank 4: Ultraspherical / Jacobi evaluation) -------------------------
 * UPSTREAM ApproxFunOrthogonalPolynomials `clenshaw(sp::PolynomialSpace, c, x)` (call sites of the generic
 * `clenshaw(sp, c, x)`: src/Extras/sample.jl:47).  p_{j+1} = (A_j x + B_j) p_j - C_j p_{j-1}, p_0 = 1:
 *   bk1 = bk2 = 0;  for j = N-1 .. 1:  bk2, bk1 = bk1, muladd(muladd(A_j,x,B_j), bk1, muladd(-C_{j+1}, bk2, c_j))
 *   return muladd(muladd(A_0,x,B_0), bk1, muladd(-C_1, bk2, c_0)).   A, B: N entries; C: N+1 entries (C_0 unused). */
void orc_clenshaw_rec(const double* c, int64_t N, const double* A, const double* B, const double* C, const double* x,
                      double* y, int64_t P, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < P; ++i) {
        if (N == 0) { y[i] = 0.0; continue; }
        double bk1 = 0.0, bk2 = 0.0;
        for (int64_t j = N - 1; j >= 1; --j) {
            const double t = fma(fma(A[j], x[i], B[j]), bk1, fma(-C[j + 1], bk2, c[j]));
            bk2 = bk1;
            bk1 = t;
        }
        y[i] = fma(fma(A[0], x[i], B[0]), bk1, fma(-C[1], bk2, c[0]));
    }
}

/* ---- bisection (a14) ------------------------------------------------------------------------ */
void orc_cheb_bisect(const double* c, int64_t N, const double* u, double* out, int64_t P, int numits,
                     int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < P; ++i) {
        double a = -1.0, b = 1.0;
        for (int it = 0; it < numits; ++it) {
            const double m = 0.5 * (a + b);
            const double v = clenshaw1(c, N, 1, m);
            if (v <= u[i]) a = m; else b = m;
        }
        out[i] = 0.5 * (a + b);
    }
}

void orc_cheb_bisect_cols(const double* C, int64_t N, int64_t ld, const double* u, double* out,
                          int64_t P, int numits, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t j = 0; j < P; ++j) {
        double a = -1.0, b = 1.0;
        const double* c = C + j * ld;
        for (int it = 0; it < numits; ++it) {
            const double m = 0.5 * (a + b);
            const double v = clenshaw1(c, N, 1, m);
            if (v <= u[j]) a = m; else b = m;
        }
        out[j] = 0.5 * (a + b);
    }
}

/* Reference-shaped vector bisection (sample.jl:58-75): whole-vector passes through the planned
 * Clenshaw, 47 times.  work must hold 7*P doubles (a,b,m,vals + 3 plan vectors). */
void orc_cheb_bisect_planned(const double* c, int64_t N, const double* u, double* out, int64_t P,
                             int numits, double* work, int nthreads) {
    double* a = work;
    double* b = work + P;
    double* m = work + 2 * P;
    double* vals = work + 3 * P;
    double* plan = work + 4 * P;
    for (int64_t i = 0; i < P; ++i) { a[i] = -1.0; b[i] = 1.0; }
    for (int it = 0; it < numits; ++it) {
        for (int64_t i = 0; i < P; ++i) m[i] = 0.5 * (a[i] + b[i]);
        orc_clenshaw_planned(c, N, m, vals, P, plan, nthreads);
        for (int64_t i = 0; i < P; ++i) {
            if (vals[i] <= u[i]) a[i] = m[i]; else b[i] = m[i];
        }
    }
    for (int64_t i = 0; i < P; ++i) out[i] = 0.5 * (a[i] + b[i]);
}

/* ---- chebnormalizedcumsum! (a12) ------------------------------------------------------------ */
/* One column, in place.  `N` input coefficients; if grow != 0 the column has room for N+1 and the
 * result has N+1 entries (Vector method of ultraint!), else the top term is dropped (Matrix method). */
static void normcumsum_col(double* v, int64_t N, int grow) {
    int64_t n = N;
    /* ultraconversion! */
    if (n == 2) {
        v[1] /= 2;
    } else if (n > 2) {
        v[0] -= v[2] / 2;
        for (int64_t j = 1; j <= n - 3; ++j) v[j] = (v[j] - v[j + 2]) / 2;
        v[n - 2] /= 2;
        v[n - 1] /= 2;
    }
    /* ultraint! */
    int64_t L = grow ? n + 1 : n;
    for (int64_t k = L - 1; k >= 1; --k) v[k] = v[k - 1] / (double)k;
    if (L > 0) v[0] = 0.0;
    /* subtract_zeroatleft!  sample.jl:121-127: f[1] += (-1)^k f[k], k = 2..L (1-based) */
    for (int64_t k = 2; k <= L; ++k) v[0] += ((k & 1) ? -1.0 : 1.0) * v[k - 1];
    /* multiply_oneatright!  sample.jl:137-150 */
    double val = 0.0;
    for (int64_t k = 0; k < L; ++k) val += v[k];
    val = 1 / val;
    for (int64_t k = 0; k < L; ++k) v[k] *= val;
}

void orc_cheb_normcumsum(double* C, int64_t N, int64_t ncols, int64_t ld, int grow) {
    for (int64_t j = 0; j < ncols; ++j) normcumsum_col(C + j * ld, N, grow);
}

/* sample(f::LowRankFun{Chebyshev,Chebyshev}, n), src/Extras/sample.jl:208-213, given ry and the uniforms:
 *   fA = evaluate.(f.A, ry') ; AB = CB*fA ; chebnormalizedcumsum!(AB) ; rx = chebbisectioninv(AB, u) ; fromcanonical.
 * BLAS leaves the summation order of CB*fA unspecified; this restatement accumulates over r in order with fma
 * (the order the GPU kernel uses) so that the two can be compared bit for bit.  A: NA x R, CB: N x R column-major. */
void orc_sample2d_lowrank(const double* A, int64_t NA, const double* CB, int64_t N, int64_t R, double ya, double yb,
                          double xa, double xb, const double* ry, const double* u, double* rx, int64_t P, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t j = 0; j < P; ++j) {
        double* col = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
        const double y = ry[j];
        const double lo_d = ya < yb ? ya : yb, hi_d = ya < yb ? yb : ya;
        const int inside = (y >= lo_d) && (y <= hi_d);
        const double yc = (2.0 * y - ya - yb) / (yb - ya);
        for (int64_t k = 0; k < N; ++k) col[k] = 0.0;
        for (int64_t r = 0; r < R; ++r) {
            const double fa = inside ? clenshaw1(A + r * NA, NA, 1, yc) : 0.0;
            for (int64_t k = 0; k < N; ++k) col[k] = fma(CB[r * N + k], fa, col[k]);
        }
        normcumsum_col(col, N, 0);
        double a = -1.0, b = 1.0;
        for (int it = 0; it < 47; ++it) {
            const double m = 0.5 * (a + b);
            if (clenshaw1(col, N, 1, m) <= u[j]) a = m; else b = m;
        }
        rx[j] = (xa + xb) * 0.5 + ((xb - xa) * (0.5 * (a + b))) * 0.5;
        free(col);
    }
}

/* bivariate evaluation in ProductFun order: g_j(x) per column, then the recurrence over j in y */
void orc_clenshaw2d(const double* C, int64_t n1, int64_t n2, const double* x, const double* y, double* out, int64_t P,
                    int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t p = 0; p < P; ++p) {
        if (n2 == 0) { out[p] = 0.0; continue; }
        const double y2 = 2.0 * y[p];
        double B1 = 0.0, B2 = 0.0;
        for (int64_t j = n2 - 1; j >= 1; --j) {
            const double g = clenshaw1(C + j * n1, n1, 1, x[p]);
            const double t = fma(y2, B1, g - B2);
            B2 = B1;
            B1 = t;
        }
        const double g0 = clenshaw1(C, n1, 1, x[p]);
        out[p] = (n2 == 1) ? g0 : fma(y2 / 2.0, B1, g0 - B2);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
