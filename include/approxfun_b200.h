/*
 * libapproxfun_b200.so -- C ABI of the B200-native ApproxFun transform / evaluate / sample path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / CUDA types), returns
 * an int32 status (AFB_OK == 0) and never throws, exits or falls back to the CPU: without a usable
 * CUDA device every compute call returns AFB_NO_DEVICE.  The message of the last failure on the
 * calling thread is available through afb_last_error().
 *
 * "Replaces" lines cite the reference interface each entry point stands in for.  Paths are relative
 * to /root/reference (ApproxFun.jl v0.13.28); UPSTREAM = package not vendored there (see SURVEY.md
 * section 8b).  Arrays follow Julia's column-major layout: a batch of B transforms of length n is an
 * n x B matrix, i.e. B contiguous vectors (`stride` elements apart).
 *
 * Ownership: the caller owns every buffer it passes; the library owns device workspaces, twiddle
 * tables and streams inside the opaque plan.  Host-pointer entry points copy in, compute on the GPU
 * and copy out before returning.  `*_device` entry points take device pointers plus a CUDA stream
 * handle (cudaStream_t passed as void*, NULL = legacy default stream) and are asynchronous.
 *
 * Threading: plan creation / destruction and table construction are mutex protected; executing
 * different plans concurrently (from different threads) is allowed -- every host-pointer call leases its own
 * streams and staging buffers from a per-device pool; a plan executed concurrently with itself is serialised
 * by a per-plan mutex.
 */
#ifndef APPROXFUN_B200_H
#define APPROXFUN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFB_VERSION 100 /* 0.1.0 */

enum afb_status {
    AFB_OK = 0,
    AFB_BAD_ARG = 1,     /* null pointer, unknown kind ...            (Julia: ArgumentError)      */
    AFB_BAD_SIZE = 2,    /* negative / inconsistent sizes             (Julia: DimensionMismatch,
                            the `@assert size(c)[2] == length(xl)` of src/Extras/sample.jl:85)  */
    AFB_NO_DEVICE = 3,   /* no CUDA device: there is NO CPU fallback                              */
    AFB_CUDA_ERROR = 4,
    AFB_NCCL_ERROR = 5,
    AFB_OOM = 6,
    AFB_UNSUPPORTED = 7
};

/* Transform kinds.  FWD = values -> coefficients (plan_transform), INV = coefficients -> values
 * (plan_itransform). */
enum afb_kind {
    AFB_CHEB1_FWD = 0,   /* plan_transform(Chebyshev(), n):  UPSTREAM FastTransforms
                            plan_chebyshevtransform(v, Val(1)) = FFTW REDFT10, /n, c0/2
                            (boundary visible in-tree: ext/ApproxFunDualNumbersExt.jl:10-12,44,49) */
    AFB_CHEB1_INV = 1,   /* plan_itransform(Chebyshev(), n): plan_ichebyshevtransform(c, Val(1)) =
                            c0*2, FFTW REDFT01, /2  (ext/ApproxFunDualNumbersExt.jl:45,50)         */
    AFB_FOURIER_FWD = 2, /* plan_transform(Fourier(), n): R2HC + interlace; in-tree twin
                            src/Extras/fftGeneric.jl:47-54, wrappers ext/...:53-63                */
    AFB_FOURIER_INV = 3, /* plan_itransform(Fourier(), n): src/Extras/fftGeneric.jl:56-64          */
    AFB_LAURENT_FWD = 4, /* plan_transform(Laurent(), n): c2c fft /n, order [f0,f-1,f1,f-2,...]
                            (docs/src/usage/spaces.md:113-125; plan_fft import src/ApproxFun.jl:37) */
    AFB_LAURENT_INV = 5,
    AFB_CHEB1_2D_FWD = 6, /* transform!(::TensorSpace{Chebyshev,Chebyshev}, M): UPSTREAM
                             ApproxFunBase; callers src/Plot/Plot.jl:290-332, src/Extras/sample.jl:196 */
    AFB_CHEB1_2D_INV = 7,
    /* "next" rows of SURVEY.md section 8f (sibling spaces on the same FFT core; correctness path, not tuned): */
    AFB_CHEB2_FWD = 8,    /* second-kind points x_j = cos(pi j/(n-1)): UPSTREAM FastTransforms
                             plan_chebyshevtransform(v, Val(2)) = FFTW REDFT00, /(n-1), ends halved
                             (Val(kind) boundary: ext/ApproxFunDualNumbersExt.jl:44-45)                      */
    AFB_CHEB2_INV = 9,    /* plan_ichebyshevtransform(c, Val(2)): ends doubled, REDFT00, /2                 */
    AFB_SIN_FWD = 10,     /* plan_transform(SinSpace(), n): src/Extras/fftGeneric.jl:74-78 (in-tree spec):
                             imag(fft([0;-x;0;reverse(x)]))[2:n+1] / (n+1)                                  */
    AFB_SIN_INV = 11,     /* plan_itransform(SinSpace(), n): src/Extras/fftGeneric.jl:80-81: same / 2       */
    AFB_TAYLOR_FWD = 12,  /* Taylor = Hardy{true}: f(t) = sum_k f_k e^{ikt} (docs/src/usage/spaces.md:100-104):
                             f_k = F_k / n on t_j = 2 pi j / n; complex in/out                              */
    AFB_TAYLOR_INV = 13,
    AFB_HARDYM_FWD = 14,  /* Hardy{false}: f(t) = sum_k f_k e^{-i(k+1)t} (spaces.md:106-110): f_k = F_{n-1-k}/n */
    AFB_HARDYM_INV = 15,
    AFB_PADUA_FWD = 16,   /* plan_transform(Chebyshev()^2, v::Vector) = FastTransforms.plan_paduatransform!(v, Val{false})
                             (UPSTREAM; NEWS.md:85): N = (d+1)(d+2)/2 values at afb_paduapoints -> N coefficients in the
                             order T0T0, T0(x)T1(y), T1(x)T0(y), ... (docs/src/usage/constructors.md:98-116).
                             afb_plan_create(kind, n = N, 0, 1, N, 0); AFB_BAD_SIZE unless N is triangular.        */
    AFB_PADUA_INV = 17,   /* plan_itransform(Chebyshev()^2, c::Vector) = plan_ipaduatransform!: coefficients -> values */
    /* mixed tensor spaces (UPSTREAM ApproxFunBase transform!(::TensorSpace, M): factor-1 plan on every column, factor-2 plan
     * on every row): PeriodicSegment x Interval as in test/PeriodicXIntervalTest.jl:22-24, test/LaplaceInAStripTest.jl:6-8 */
    AFB_FOURIER_CHEB1_2D_FWD = 18, /* Fourier (x) Chebyshev: real n x n2 matrix                                   */
    AFB_FOURIER_CHEB1_2D_INV = 19,
    AFB_LAURENT_CHEB1_2D_FWD = 20, /* Laurent (x) Chebyshev: complex n x n2 matrix (interleaved complex double)   */
    AFB_LAURENT_CHEB1_2D_INV = 21
};

typedef struct afb_plan_s* afb_plan;

/* ---- library ---------------------------------------------------------------------------------- */
int32_t afb_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its status code */
int32_t afb_last_error(char* buf, size_t len);
/* Select the CUDA devices of this process.
 *   ndev == 0 (devices may be NULL): keep the calling thread's current device.
 *   ndev == 1: make devices[0] current (the one-process-per-GPU layout of torchrun / MPI launches).
 *   ndev  > 1: ONE process drives several GPUs (what a Julia session does): devices[0] becomes current, peer access is
 *              enabled between all pairs, and the host-pointer entry points fan out over the set with one internal host
 *              thread and one stream set per device, joined before the call returns -- afb_plan_execute_host shards the
 *              batch of a 1-D plan by contiguous blocks of functions and runs a 2-D plan as column slabs -> peer-memory
 *              transpose over NVLink -> row slabs; afb_clenshaw_cheb / afb_clenshaw_rec / afb_cheb_bisect /
 *              afb_samplecdf_cheb / afb_sample_cheb shard their point sets.  Results are bit-identical to one device.
 * At most 16 devices (ordinals 0..15, one node).  Plans, tables and staging buffers belong to the device that was current
 * when they were created; a plan may be executed from any thread.  *_device entry points always run on ONE device (the
 * plan's, or the calling thread's current device for the plain functions). */
int32_t afb_init(const int32_t* devices, int32_t ndev);
/* Frees the staging buffers and cached device blocks of every device.  Plans stay valid. */
int32_t afb_shutdown(void);
int32_t afb_device_info(int32_t* sm_count, int64_t* hbm_bytes, char* name, size_t name_len);

/* ---- device / pinned memory helpers (for device-resident use and benchmarking) ---------------- */
int32_t afb_malloc(void** dptr, size_t bytes);
int32_t afb_free(void* dptr);
int32_t afb_host_alloc(void** hptr, size_t bytes); /* page-locked */
int32_t afb_host_free(void* hptr);
/* Page-locks a caller-owned host array in place (cudaHostRegister, visible to every device) so that the host entry points
 * DMA it directly instead of staging it through the library's pinned ring (afb_plan_execute_host on a pageable
 * 65 536 x 4 096 array: 110-137 ms; registered or afb_host_alloc'ed: 45-58 ms).  The caller MUST call afb_host_unregister
 * before the array is freed (in Julia: pin for the lifetime of a work array, unpin in its finalizer); the library never
 * registers caller memory on its own, because it cannot see the array die. */
int32_t afb_host_register(void* hptr, size_t bytes);
int32_t afb_host_unregister(void* hptr);
int32_t afb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int32_t afb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int32_t afb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int32_t afb_stream_sync(void* stream);

/* ---- transform plans ----------------------------------------------------------------------------
 * Replaces: plan_transform / plan_itransform (import src/ApproxFun.jl:22-23) and the FFTW plan they
 * wrap (UPSTREAM FFTW.plan_r2r / plan_fft); `plan * v` (NEWS.md:90) = afb_plan_execute_host.
 *   n      transform length (1D) or number of rows (2D)
 *   n2     0 for 1D kinds; number of columns for 2D kinds
 *   batch  number of independent transforms (1D); must be 1 for 2D
 *   stride elements (double, or complex double for Laurent) between consecutive transforms; >= n
 *   flags  0, or AFB_PLAN_2D_TRANSPOSE: 2D kinds run the row pass as transpose + contiguous lines + transpose
 *          (the only route when n is odd or n2 is not a power of two in 2^10..2^18) instead of the strided
 *          four-step on row pairs; results agree to rounding.
 * Data type: double for Chebyshev / Fourier, interleaved complex double for Laurent.
 * Any n >= 0 is accepted, as FFTW accepts any length: n <= 32 a direct kernel; powers of two one shared-memory line kernel
 * (n <= 16384) or a four-step; n = 2^a 3^b 5^c 7^d <= 6912 one mixed-radix kernel; other n <= 4096 one fused Bluestein kernel;
 * n = S h (S = 2, 3, 4, 5, 7, 8, 16) with h inside one of those engines three launches; remaining n Bluestein over a padded copy. */
#define AFB_PLAN_2D_TRANSPOSE 1u
int32_t afb_plan_create(afb_plan* out, int32_t kind, int64_t n, int64_t n2, int64_t batch, int64_t stride,
                        uint32_t flags);
/* host buffers; in == out allowed.  H2D, kernels and D2H are pipelined over batch chunks. */
int32_t afb_plan_execute_host(afb_plan plan, const void* in, void* out);
/* device buffers; in == out allowed; asynchronous on `stream` */
int32_t afb_plan_execute_device(afb_plan plan, const void* d_in, void* d_out, void* stream);
int32_t afb_plan_destroy(afb_plan plan);
/* doubles a plan reads / writes per execution (complex kinds count two per element; strided batches include the gaps) */
int32_t afb_plan_io_doubles(afb_plan plan, int64_t* in_doubles, int64_t* out_doubles);
/* number of kernels one afb_plan_execute_device call launches (bench.py gpu_launches) */
int32_t afb_plan_launches(afb_plan plan, int32_t* launches);

/* ---- points ---------------------------------------------------------------------------------------
 * Replaces points(::Chebyshev, n): UPSTREAM FastTransforms.chebyshevpoints(Float64,n,Val(1)) mapped by
 * fromcanonical; DESCENDING order (NEWS.md:92).  out[k] = (a+b)/2 + (b-a)*sinpi((n-2k-1)/(2n))/2. */
int32_t afb_chebpoints(double* out, int64_t n, double a, double b);
int32_t afb_chebpoints_device(double* d_out, int64_t n, double a, double b, void* stream);
/* points(::Fourier/Laurent, n): out[j] = a + (b-a) j / n */
int32_t afb_fourierpoints(double* out, int64_t n, double a, double b);
/* points(Chebyshev(ax..bx) x Chebyshev(ay..by), N): UPSTREAM FastTransforms.paduapoints(d), d = cld(-3+sqrt(1+8N),2),
 * mapped by fromcanonical.  xy is Np x 2 column-major (x then y), Np = (d+1)(d+2)/2 >= N; *np_out receives Np.
 * Pass xy == NULL to query Np.  Emission order: k = d..0 (x = -cospi(k/d)), j = NN+delta..1. */
int32_t afb_paduapoints(double* xy, int64_t N, double ax, double bx, double ay, double by, int64_t* np_out);

/* ---- Clenshaw evaluation ----------------------------------------------------------------------------
 * Replaces clenshaw(::Chebyshev, c, x) and clenshaw(c::Vector, x::Vector, plan::ClenshawPlan) (UPSTREAM
 * ApproxFunOrthogonalPolynomials; import src/ApproxFun.jl:26; call sites src/Extras/sample.jl:47,56,68).
 * y[i] = sum_k c[k] T_k(x[i]); x is NOT modified and y never aliases internal scratch. */
int32_t afb_clenshaw_cheb(const double* c, int64_t N, const double* x, double* y, int64_t P);
int32_t afb_clenshaw_cheb_device(const double* d_c, int64_t N, const double* d_x, double* d_y, int64_t P,
                                 void* stream);
/* General three-term Clenshaw: y[i] = sum_j c[j] p_j(x[i]) with p_0 = 1,
 * p_{j+1} = (A[j] x + B[j]) p_j - C[j] p_{j-1}  (recA / recB / recC of UPSTREAM ApproxFunOrthogonalPolynomials: Jacobi,
 * Ultraspherical, Legendre; DLMF 18.9.1).  A, B: N entries, C: N+1 entries (C[0] unused).  Operation order of UPSTREAM
 * `clenshaw(sp::PolynomialSpace, c, x)`: bk1 <- muladd(muladd(A_j,x,B_j), bk1, muladd(-C_{j+1}, bk2, c_j)), j = N-1..0. */
int32_t afb_clenshaw_rec(const double* c, int64_t N, const double* A, const double* B, const double* C, const double* x,
                         double* y, int64_t P);
int32_t afb_clenshaw_rec_device(const double* d_c, int64_t N, const double* d_A, const double* d_B, const double* d_C,
                                const double* d_x, double* d_y, int64_t P, void* stream);
/* Replaces clenshaw(c::Matrix, x::Vector, plan) (call sites src/Extras/sample.jl:80,94): column j of the
 * N x P column-major matrix C (leading dimension ld >= N) evaluated at x[j]. */
int32_t afb_clenshaw_cheb_cols(const double* C, int64_t N, int64_t ld, const double* x, double* y, int64_t P);
int32_t afb_clenshaw_cheb_cols_device(const double* d_C, int64_t N, int64_t ld, const double* d_x, double* d_y,
                                      int64_t P, void* stream);
/* Replaces evaluate.(f.A, ry') (src/Extras/sample.jl:208): each of the M columns of C (N x M column-major)
 * at each of the P points; Y is M x P column-major, Y[m + M*i] = f_m(x[i]). */
int32_t afb_clenshaw_cheb_multi(const double* C, int64_t N, int64_t M, const double* x, double* Y, int64_t P);
int32_t afb_clenshaw_cheb_multi_device(const double* d_C, int64_t N, int64_t M, const double* d_x, double* d_Y,
                                       int64_t P, void* stream);

/* ---- sampling -----------------------------------------------------------------------------------------
 * Replaces chebnormalizedcumsum!(f) (src/Extras/sample.jl:170-175 = ultraconversion!, ultraint!,
 * subtract_zeroatleft!, multiply_oneatright!), in place on ncols columns of length N (leading dimension
 * ld).  grow != 0: Vector semantics, each column gains one coefficient (needs ld >= N+1);
 * grow == 0: Matrix semantics, the top term is dropped. */
int32_t afb_cheb_normcumsum(double* C, int64_t N, int64_t ncols, int64_t ld, int32_t grow);
int32_t afb_cheb_normcumsum_device(double* d_C, int64_t N, int64_t ncols, int64_t ld, int32_t grow, void* stream);
/* Replaces chebbisectioninv(c::Vector, u::Vector[, plan]) (src/Extras/sample.jl:55-75): numits (47)
 * Clenshaw bisections on [-1,1] per entry of u; returns the final midpoint (canonical variable). */
int32_t afb_cheb_bisect(const double* c, int64_t N, const double* u, double* out, int64_t P, int32_t numits);
int32_t afb_cheb_bisect_device(const double* d_c, int64_t N, const double* d_u, double* d_out, int64_t P,
                               int32_t numits, double a, double b, void* stream);
/* Replaces chebbisectioninv(c::Matrix, u::Vector[, plan]) (src/Extras/sample.jl:79-101). */
int32_t afb_cheb_bisect_cols(const double* C, int64_t N, int64_t ld, const double* u, double* out, int64_t P,
                             int32_t numits);
int32_t afb_cheb_bisect_cols_device(const double* d_C, int64_t N, int64_t ld, const double* d_u, double* d_out,
                                    int64_t P, int32_t numits, void* stream);
/* Replaces bisectioninv(cf::Fun{<:Chebyshev,Float64}, u::Vector) (src/Extras/sample.jl:103-109):
 * fromcanonical.(a..b, chebbisectioninv(cdf_c, u)) with caller-supplied uniforms (`samplecdf`,
 * src/Extras/sample.jl:184 with rand(n) provided by the caller). */
int32_t afb_samplecdf_cheb(const double* cdf_c, int64_t N, double a, double b, const double* u, double* out,
                           int64_t P);
/* Throughput variant of sample(f, n) (src/Extras/sample.jl:182): uniforms come from an on-device
 * counter-based generator (Philox4x32-10, stream `seed`, counter offset `offset`); output on device. */
/* sample(f, n) with the uniforms drawn on the device (host output): sample i uses Philox4x32-10 counter i of `seed`, so the
 * stream is reproducible and identical to afb_sample_cheb_device(..., seed, 0, ...).  cdf_c = normalizedcumsum coefficients. */
int32_t afb_sample_cheb(const double* cdf_c, int64_t N, double a, double b, uint64_t seed, double* out, int64_t P);
int32_t afb_sample_cheb_device(const double* d_cdf_c, int64_t N, double a, double b, uint64_t seed,
                               uint64_t offset, double* d_out, int64_t P, void* stream);
/* the uniforms afb_sample_cheb_device draws, for parity checks */
int32_t afb_philox_uniform_device(uint64_t seed, uint64_t offset, double* d_out, int64_t P, void* stream);

/* Piecewise Chebyshev Funs (a Fun on a PiecewiseSpace of Chebyshev segments, e.g. abs(Fun(sin, -5..5))): piece k lives on
 * breaks[k]..breaks[k+1] (increasing) with coefficients coef[offsets[k] .. offsets[k+1]).
 * afb_piecewise_clenshaw: y[i] = f(x[i]) in the domain variable -- the FIRST piece whose closed interval contains x[i],
 *   evaluated at tocanonical = (2x - lo - hi)/(hi - lo); zero outside the domain (replaces evaluate(f::Fun{<:PiecewiseSpace}, x)).
 * afb_piecewise_bisect: the GENERIC bisectioninv(f::Fun, x; numits = 47) of src/Extras/sample.jl:23-36 with genmidpoint
 *   (sample.jl:11-21): a, b = minimum / maximum of the domain; numits x { m = genmidpoint(a,b); f(m) <= u ? a = m : b = m };
 *   returns .5(a+b) per entry of u.  This is the route sample(f, n) takes for piecewise Funs (test/PiecewiseSampleTest.jl:5-10);
 *   with npieces == 1 it is the generic bisection of a plain Chebyshev Fun. */
int32_t afb_piecewise_clenshaw(const double* coef, const int64_t* offsets, const double* breaks, int64_t npieces, const double* x,
                               double* y, int64_t P);
int32_t afb_piecewise_bisect(const double* coef, const int64_t* offsets, const double* breaks, int64_t npieces, const double* u,
                             double* out, int64_t P, int32_t numits);

/* Bivariate evaluation f(x_p, y_p) = sum_{i,j} C[i,j] T_i(x_p) T_j(y_p) of a Chebyshev (x) Chebyshev coefficient matrix
 * (n1 x n2 column-major; canonical variables), in the order a ProductFun is evaluated upstream (columns in x by the
 * 1-D recurrence, then the recurrence over columns in y).  Replaces f(x, y) / evaluate for TensorSpace / ProductFun Funs
 * (callers: examples/PDE_Helmholtz.jl:33-36, src/Plot/Plot.jl:335-338, test/PeriodicXIntervalTest.jl:22-24). */
int32_t afb_clenshaw_cheb2d(const double* C, int64_t n1, int64_t n2, const double* x, const double* y, double* out,
                            int64_t P);
int32_t afb_clenshaw_cheb2d_device(const double* d_C, int64_t n1, int64_t n2, const double* d_x, const double* d_y,
                                   double* d_out, int64_t P, void* stream);

/* The same for a Fourier (x) Chebyshev coefficient matrix (rows in the Fourier order [1, sin t, cos t, sin 2t, ...],
 * docs/src/usage/spaces.md:86-98): f(t_p, y_p) = sum_{i,j} C[i,j] F_i(t_p) T_j(y_p), t canonical in [0, 2 pi), y in [-1, 1].
 * Replaces f(x, y) for Funs on PeriodicSegment x Interval (test/PeriodicXIntervalTest.jl:22-24). */
int32_t afb_clenshaw_fourier_cheb2d(const double* C, int64_t n1, int64_t n2, const double* t, const double* y, double* out,
                                    int64_t P);
int32_t afb_clenshaw_fourier_cheb2d_device(const double* d_C, int64_t n1, int64_t n2, const double* d_t, const double* d_y,
                                           double* d_out, int64_t P, void* stream);

/* ---- 2-D sampling (sample(f::LowRankFun{Chebyshev,Chebyshev}, n), src/Extras/sample.jl:205-214) -------------
 * Fused replacement of lines 208-213: fA = evaluate.(f.A, ry'); AB = CB*fA; chebnormalizedcumsum!(AB);
 * rx = chebbisectioninv(AB, u); fromcanonical.(domain(f,1), rx).  A is NA x R and CB is N x R (column-major
 * coefficient matrices of f.A and f.B), ya..yb the domain f.A is evaluated on, xa..xb = domain(f,1), ry the
 * samples of the other variable, u the caller's rand(n).  The N x n matrix AB is never materialised (N <= 198). */
int32_t afb_sample2d_cheb_lowrank(const double* A, int64_t NA, const double* CB, int64_t N, int64_t R, double ya,
                                  double yb, double xa, double xb, const double* ry, const double* u, double* rx,
                                  int64_t P);
int32_t afb_sample2d_cheb_lowrank_device(const double* d_A, int64_t NA, const double* d_CB, int64_t N, int64_t R,
                                         double ya, double yb, double xa, double xb, const double* d_ry,
                                         const double* d_u, double* d_rx, int64_t P, void* stream);
/* AB = CB*fA (N x R times R x P, column-major): the dgemm of src/Extras/sample.jl:210 as a stand-alone call. */
int32_t afb_dgemm_nn(const double* CB, const double* fA, double* AB, int64_t N, int64_t R, int64_t P);
int32_t afb_dgemm_nn_device(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P,
                            void* stream);

/* Measurement variant of afb_dgemm_nn_device on the FP64 tensor cores (DMMA m8n8k4), kept to answer with numbers whether the
 * one dense contraction of the path (sample.jl:210) belongs on them; same result to rounding. */
int32_t afb_dgemm_nn_dmma_device(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P,
                                 void* stream);

/* Measurement helper (no reference counterpart): runs `blocks` x 256 threads x 8 chains x `iters` FP64 steps;
 * mode 0 = DFMA with three distinct register operands, 1 = DFMA with a repeated operand, 2 = the Clenshaw DADD+DFMA step,
 * 3 = the reassociated step with a warp-uniform coefficient, 4 = DMMA m8n8k4 (8 accumulator tiles per warp, 512 flop each).
 * d_out needs blocks*256 doubles.  bench.py times it to report a MEASURED FP64 rate next to the nominal peak. */
int32_t afb_fp64_probe_device(int32_t mode, int32_t iters, int32_t blocks, double* d_out, void* stream);

/* ---- device-resident Fun algebra ------------------------------------------------------------
 * Opaque device buffers of doubles plus point-wise kernels, so that the chains the reference builds from host arrays --
 *   Fun(f, S, n):   vals = f.(points(S, n)); transform(S, vals)            (ext/ApproxFunDualNumbersExt.jl:79-82)
 *   adaptive Fun:   the ladder n = 2^4 .. 2^20, 2^21 with its tail test     (ext/ApproxFunDualNumbersExt.jl:96-111)
 *   f * g, sin(f):  values(pad(f, n)) -> point-wise op -> transform -> chop (UPSTREAM ApproxFunBase)
 * -- stay on the GPU between steps: at most one upload of samples and one download of the final coefficients.
 * All calls run on the legacy default stream of the buffer's device, in call order. */
typedef struct afb_buf_s* afb_buf;
int32_t afb_buf_create(afb_buf* out, int64_t n_doubles);
int32_t afb_buf_destroy(afb_buf buf);
int32_t afb_buf_size(afb_buf buf, int64_t* n_doubles);
int32_t afb_buf_ptr(afb_buf buf, void** dptr);     /* raw device pointer, for the *_device entry points */
int32_t afb_buf_upload(afb_buf buf, const double* host, int64_t n, int64_t offset);
int32_t afb_buf_download(afb_buf buf, double* host, int64_t n, int64_t offset);
/* plan * v with v and the result on the device (in == out allowed) */
int32_t afb_plan_execute_buf(afb_plan plan, afb_buf in, afb_buf out);
/* out[i] = op(alpha * a[i] + beta), i < n   (in place allowed) */
enum afb_unary_op { AFB_OP_IDENT = 0, AFB_OP_SIN, AFB_OP_COS, AFB_OP_EXP, AFB_OP_ABS, AFB_OP_SQRT, AFB_OP_SQUARE, AFB_OP_RECIP,
                    AFB_OP_LOG, AFB_OP_TANH };
int32_t afb_buf_unary(int32_t op, afb_buf a, afb_buf out, int64_t n, double alpha, double beta);
/* out[i] = a[i] (+, -, *, /) b[i] */
enum afb_binary_op { AFB_OP_ADD = 0, AFB_OP_SUB, AFB_OP_MUL, AFB_OP_DIV };
int32_t afb_buf_binary(int32_t op, afb_buf a, afb_buf b, afb_buf out, int64_t n);
/* out[0..n_out) = in[0..min(n_in, n_out)) then zeros: pad(f, n) / truncation of a coefficient vector */
int32_t afb_buf_pad(afb_buf in, int64_t n_in, afb_buf out, int64_t n_out);
/* points(::Chebyshev, n) into a buffer (same kernel as afb_chebpoints) */
int32_t afb_buf_chebpoints(afb_buf out, int64_t n, double a, double b);
/* y = clenshaw(c[0..N), x[0..P)) on buffers (canonical variable) */
int32_t afb_buf_clenshaw(afb_buf c, int64_t N, afb_buf x, afb_buf y, int64_t P);
/* The ladder's tail test and chop on the device: *maxabs = max |c|, *tailmax = max |c[n-ntail..n)|, and (chop_len != NULL)
 * *chop_len = 1 + the last index with |c| > rel_thresh * max |c| (0 if none).  Transfers 24 bytes. */
int32_t afb_buf_tail(afb_buf c, int64_t n, int64_t ntail, double rel_thresh, double* maxabs, double* tailmax, int64_t* chop_len);

/* ---- multi-GPU 2D transform (one process per GPU) ----------------------------------------------------
 * The n x n2 matrix is sharded by contiguous column blocks; the plan performs local column transforms,
 * an all-to-all transpose over NVLink, and local transforms of the other factor.  The caller supplies
 * an NCCL unique id (128 bytes, broadcast by the host layer) -- see DESIGN.md. */
int32_t afb_comm_unique_id(void* id128);
int32_t afb_comm_create(const void* id128, int32_t rank, int32_t world);
int32_t afb_comm_destroy(void);
/* kind = AFB_CHEB1_2D_FWD / _INV; n and n2 must be multiples of the world size.
 * d_in : this rank's column block (column-major n x n2/world)
 * d_out: this rank's block of rows of the result, as n/world contiguous lines of length n2 */
int32_t afb_sharded2d_create(void** handle, int32_t kind, int64_t n, int64_t n2);
int32_t afb_sharded2d_execute_device(void* handle, const double* d_in, double* d_out, void* stream);
int32_t afb_sharded2d_destroy(void* handle);
/* Optional peer-memory mode: replaces pack + NCCL send/recv + unpack by ONE kernel that transposes tiles straight
 * into the owning rank's buffer over NVLink (CUDA IPC mapped peer memory).  Each rank exports 64 bytes, the host layer
 * all-gathers them (world x 64 bytes) and passes the array to afb_sharded2d_ipc_open on every rank. */
/* world <= 16 (the peer-pointer tables are fixed size; one node). */
int32_t afb_sharded2d_ipc_export(void* handle, void* handle64);
int32_t afb_sharded2d_ipc_open(void* handle, const void* handles, int32_t world);
/* Peer-memory mode only: run the column pass in `chunks` (1..8) pieces and send each finished piece to the peers on a side
 * stream while the next is transformed (default 1 = one column pass, one exchange kernel). */
int32_t afb_sharded2d_set_chunks(void* handle, int32_t chunks);
/* Measurement only (peer-memory mode): run a subset of the phases -- bit 0 column pass, bit 1 exchange + arrival wait,
 * bit 2 row pass; 7 = the transform.  Every rank must use the same mask. */
int32_t afb_sharded2d_set_phases(void* handle, int32_t mask);
/* Tuning of the exchange kernel (defaults are the measured best): tile order 0 = one destination at a time, rotated by rank,
 * row tiles fastest; 1 = all destinations interleaved, rotated; 2 = row tiles fastest (unrotated); 3 = as 0 with column tiles
 * fastest (contiguous spans at the destination); 4 = as 3 with every 1 KB row of a 32 x 128 tile sent as one TMA bulk copy
 * (default).  grid_mult = CTAs per SM of its persistent grid; novec != 0 = 8-byte instead of 16-byte remote stores. */
int32_t afb_sharded2d_set_tuning(void* handle, int32_t order, int32_t grid_mult, int32_t novec);

#ifdef __cplusplus
}
#endif
#endif /* APPROXFUN_B200_H */
