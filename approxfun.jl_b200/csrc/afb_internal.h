// Internal declarations shared by the translation units of libapproxfun_b200.so.
#pragma once
#include <mutex>
#include <vector>

#include "afb_common.cuh"

namespace afb {

// line kinds of afb_transform.cu (values must match enum LineKind there)
enum { ILK_CHEB_FWD = 0, ILK_CHEB_INV = 1, ILK_FOURIER_FWD = 2, ILK_FOURIER_INV = 3, ILK_LAURENT_FWD = 4,
       ILK_LAURENT_INV = 5, ILK_CFFT = 6 };

constexpr int DIRECT_MAX = 32;      // n <= DIRECT_MAX: O(n^2) kernel
constexpr int LINE_MIN_N = 32;      // complex FFT lengths served by one CTA in shared memory
constexpr int LINE_MAX_N = 8192;
constexpr int LINE_MAXRL = 16;       // largest last-pass radix of the line kernels (size of the rotation-constant rows)

// ---- kernels / launchers --------------------------------------------------------------------------
int32_t launch_line(int kind, int N, const void* in, void* out, int64_t batch, int64_t istride, int64_t ostride,
                    int64_t oes, const cplx* tw, const cplx* aux, const cplx* K, double scale, void* stream);
int32_t line_last_pass(int N, int* RL, int* NSL);
int32_t launch_cfft_lines(int N, const cplx* in, cplx* out, int64_t nlines, int64_t istride, int64_t ies,
                          int64_t ostride, int64_t oes, const cplx* tw, bool conj_in, bool conj_out, double scale,
                          const cplx* twA, const cplx* twB, int64_t twmod, void* stream, int io_mode = 0,
                          int64_t bigN = 0, int64_t lpb = 0, int64_t ibs = 0, int64_t obs = 0);
int32_t launch_direct(int kind, int n, const void* in, void* out, int64_t batch, int64_t istride, int64_t ostride,
                      int64_t oes, const cplx* tab, void* stream);

int32_t launch_clenshaw_vec(const double* d_c, int64_t N, const double* d_x, double* d_y, int64_t P, void* stream);
int32_t launch_clenshaw_rec(const double* d_c, int64_t N, const double* d_A, const double* d_B, const double* d_C,
                            const double* d_x, double* d_y, int64_t P, void* stream);
int32_t launch_bisect_vec(const double* d_c, int64_t N, const double* d_u, bool rng, uint64_t seed, uint64_t offset,
                          double* d_out, int64_t P, int numits, double a, double b, void* stream);
int32_t launch_piecewise(const double* d_coef, const int32_t* d_offs, const double* d_brk, int64_t np, int64_t total,
                         const double* d_xu, double* d_out, int64_t P, bool bisect, int numits, void* stream);
int32_t launch_philox(uint64_t seed, uint64_t offset, double* d_out, int64_t P, void* stream);
int32_t launch_cols(bool bisect, const double* d_C, int64_t N, int64_t ld, const double* d_xu, double* d_out,
                    int64_t P, int numits, void* stream);
int32_t launch_multi(const double* d_C, int64_t N, int64_t M, const double* d_x, double* d_Y, int64_t P, void* stream);
int32_t launch_normcumsum(double* d_C, int64_t N, int64_t ncols, int64_t ld, int grow, void* stream);
int32_t launch_chebpoints(double* d_out, int64_t n, double a, double b, void* stream);
int32_t launch_sample2d(const double* d_A, int64_t NA, const double* d_CB, int64_t N, int64_t R, double ya, double yb,
                        double xa, double xb, const double* d_ry, const double* d_u, double* d_rx, int64_t P, int numits,
                        void* stream);
int32_t launch_clenshaw2d(const double* d_C, int64_t n1, int64_t n2, const double* d_x, const double* d_y, double* d_out,
                          int64_t P, void* stream);
int32_t launch_fourier_cheb2d(const double* d_C, int64_t n1, int64_t n2, const double* d_t, const double* d_y, double* d_out,
                              int64_t P, void* stream);
int32_t launch_fp64_probe(int mode, int iters, int blocks, double* d_out, void* stream);
int32_t launch_gemm_nn_dmma(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P, void* stream);
int32_t launch_gemm_nn(const double* d_CB, const double* d_fA, double* d_AB, int64_t N, int64_t R, int64_t P, void* stream);

// ---- generic complex FFT engine (afb_large.cu): any length, batched, contiguous lines ---------------
struct CfftPlan;
int32_t cfft_create(CfftPlan** out, int64_t n, int64_t max_batch);
void cfft_destroy(CfftPlan* p);
// out[b*n + k] = scale * sum_j in[b*n + j] exp(-/+ 2 pi i jk/n)   (inverse = +, unnormalised); in == out allowed
int32_t cfft_exec(CfftPlan* p, const cplx* in, cplx* out, int64_t batch, bool inverse, double scale, void* stream);
int32_t cfft_launches(const CfftPlan* p);
// split plans (n = S h with h inside a one-CTA engine): S, else 1.  The pre / post launchers below take the plan and fuse
// the radix-S pass / the de-interleave; cfft_exec_work runs the transform(s) of a work buffer they filled, in place.
int cfft_split(const CfftPlan* p);
int32_t cfft_exec_work(CfftPlan* p, cplx* work, int64_t nfft, bool inverse, void* stream);
int32_t cfft_work_launches(const CfftPlan* p);
// arbitrary-length real kinds in one kernel (Bluestein, padded length within one CTA's shared memory)
bool cfft_real_fused_ok(const CfftPlan* p);
int32_t launch_blue_real(CfftPlan* p, int kind, const double* in, int64_t istride, double* out, int64_t ostride, int64_t batch,
                         const cplx* qtab, void* stream);
// mixed-radix engine (afb_mixed.cu): n = 2^a 3^b 5^c 7^d, not a power of two, 6 <= n <= 6912
bool mixed_radix_ok(int64_t n);
int32_t launch_mixed_cfft(int64_t n, const cplx* tw, const cplx* in, cplx* out, int64_t batch, bool inverse, double scale,
                          void* stream);
int32_t launch_mixed_real(int kind, int64_t n, const cplx* tw, const cplx* qtab, const double* in, int64_t istride, double* out,
                          int64_t ostride, int64_t batch, void* stream);
// four-step only: real DCT-ordered input (io_mode bit 0) / output (bit 1); see LineParams::io_mode
int32_t cfft_exec_dctorder(CfftPlan* p, const void* in, void* out, bool conj_in, bool conj_out, double scale, int io_mode,
                           void* stream);
// large power-of-two real transforms: half-length four-step FFT + one pair kernel (afb_large.cu)
int32_t launch_large_post(int kind, int64_t n, const cplx* Z, double* out, void* stream);
int32_t launch_large_pre(int kind, int64_t n, const double* in, cplx* Zc, void* stream);

// extension kinds in one kernel (FFT length a power of two <= 8192, or Bluestein padded length <= 8192)
bool cfft_ext_fused_ok(const CfftPlan* p);
int32_t launch_ext_fused(CfftPlan* p, int kind, int64_t n, const double* in, int64_t istride, double* out, int64_t ostride,
                         int64_t batch, void* stream);

// pre/post kernels mapping the real transforms onto a length-n complex FFT (afb_large.cu)
int32_t launch_generic_pre(int kind, int64_t n, const void* in, int64_t istride, cplx* work, int64_t batch,
                           const cplx* qtab, void* stream, const CfftPlan* sp = nullptr);
int32_t launch_generic_post(int kind, int64_t n, const cplx* work, void* out, int64_t ostride, int64_t batch,
                            const cplx* qtab, void* stream, const CfftPlan* sp = nullptr);

int32_t launch_ext_pre(int kind, int64_t n, int64_t cn, const double* in, int64_t istride, cplx* work, int64_t batch,
                       void* stream, const CfftPlan* sp = nullptr);
int32_t launch_ext_post(int kind, int64_t n, int64_t cn, const cplx* work, double* out, int64_t ostride, int64_t batch,
                        void* stream, const CfftPlan* sp = nullptr);

// large caller arrays to / from the current device (afb_api.cu): pageable memory goes through the pinned ring
bool host_is_pageable(const void* p);
int32_t host_big_h2d(void* d, const void* h, size_t bytes);
int32_t host_big_d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t height);

// ---- 2-D (afb_tensor.cu) -----------------------------------------------------------------------------
int32_t launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, void* stream);

// row pass without transposes (afb_rows.cu): Chebyshev transform along dimension 2 of a column-major n1 x n2 matrix
bool rows_cheb_supported(int64_t n1, int64_t n2);
int64_t rows_tmp_doubles(int64_t n1, int64_t n2);  // size of `tmp` below
int32_t launch_rows_cheb(bool inverse, const double* in, double* tmp, double* out, int64_t n1, int64_t n2, void* stream);

// Padua helpers (afb_tensor.cu): dst[idx[p]] = src[p] ; dst[p] = scale * src[idx[p]] ; the point coordinates
int32_t launch_scatter_idx(double* dst, const double* src, const int32_t* idx, int64_t N, void* stream);
int32_t launch_gather_idx(double* dst, const double* src, const int32_t* idx, int64_t N, double scale, void* stream);
int32_t launch_paduapoints(double* xy, const int32_t* idx, int64_t N, int64_t d, double ax, double bx, double ay, double by,
                           void* stream);

// in-process multi-device 2-D plan (afb_tensor.cu), reached through afb_plan_execute_host
bool mdev2d_supported(int32_t kind, int64_t n, int64_t n2, int ndev, uint32_t flags);
int32_t mdev2d_create(void** out, int32_t kind, int64_t n, int64_t n2, const std::vector<int>& devs);
int32_t mdev2d_execute_host(void* handle, const double* in, double* out);
void mdev2d_destroy(void* handle);
bool mdev2d_matches(void* handle, const std::vector<int>& devs);
// `batch` lines of a 1-D plan (batch <= the plan's batch), on the plan's device; used by the chunked column pass
int32_t slab_col_chunk(afb_plan plan, const double* d_in, double* d_out, int64_t batch, void* stream);
const std::vector<int>& device_list();
// recycled device blocks (per device, freed by afb_shutdown)
int32_t device_block_get(size_t bytes, void** out, size_t* cap);
void device_block_put(void* p, size_t cap, int dev);

// ---- device table helpers ------------------------------------------------------------------------------
int32_t upload_table(const std::vector<cplx>& host, cplx** dev);

}  // namespace afb
