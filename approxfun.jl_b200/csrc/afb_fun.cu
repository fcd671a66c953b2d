// Device-resident Fun algebra: opaque device buffers, point-wise kernels and the two small
// reductions the adaptive constructor needs, so that `values -> point-wise op -> transform` chains and the ladder
// n = 2^4 .. 2^20, 2^21 (/root/reference/ext/ApproxFunDualNumbersExt.jl:96-111; UPSTREAM ApproxFunBase default_Fun) run
// without a PCIe round trip per step: samples are produced ON the device from afb_buf_chebpoints, every rung of the ladder
// is one point-wise chain + one transform, and the host only sees 16 bytes per rung (max |c| and the tail max) until the
// final coefficients are downloaded.
#include <algorithm>

#include "afb_internal.h"

struct afb_buf_s {
    double* p = nullptr;
    int64_t n = 0;     // capacity in doubles
    int dev = 0;
    size_t cap = 0;    // bytes of the underlying (recycled) device block
};

namespace afb {

enum { UOP_IDENT = 0, UOP_SIN, UOP_COS, UOP_EXP, UOP_ABS, UOP_SQRT, UOP_SQUARE, UOP_RECIP, UOP_LOG, UOP_TANH, UOP_COUNT };
enum { BOP_ADD = 0, BOP_SUB, BOP_MUL, BOP_DIV, BOP_COUNT };

template <int OP> __device__ __forceinline__ double uop(double v) {
    if (OP == UOP_SIN) return sin(v);
    if (OP == UOP_COS) return cos(v);
    if (OP == UOP_EXP) return exp(v);
    if (OP == UOP_ABS) return fabs(v);
    if (OP == UOP_SQRT) return sqrt(v);
    if (OP == UOP_SQUARE) return __dmul_rn(v, v);
    if (OP == UOP_RECIP) return __ddiv_rn(1.0, v);
    if (OP == UOP_LOG) return log(v);
    if (OP == UOP_TANH) return tanh(v);
    return v;
}
// out[i] = op(alpha * a[i] + beta): the affine part covers scalar multiples, shifts and the domain maps
template <int OP>
__global__ void __launch_bounds__(256) unary_kernel(const double* __restrict__ a, double* __restrict__ out, int64_t n, double alpha,
                                                    double beta) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = uop<OP>(__fma_rn(alpha, a[i], beta));
}
template <int OP>
__global__ void __launch_bounds__(256) binary_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                     double* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = a[i], y = b[i];
        out[i] = (OP == BOP_ADD) ? __dadd_rn(x, y) : (OP == BOP_SUB) ? __dsub_rn(x, y) : (OP == BOP_MUL) ? __dmul_rn(x, y) : __ddiv_rn(x, y);
    }
}
// out[0 .. n_out) = in[0 .. min(n_in, n_out)) followed by zeros (pad / truncate a coefficient vector)
__global__ void __launch_bounds__(256) pad_kernel(const double* __restrict__ in, int64_t n_in, double* __restrict__ out, int64_t n_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n_in ? in[i] : 0.0;
}
// res[0] = max |c|, res[1] = max |c[n-ntail .. n)|, res[2] = 1 + last index with |c| > thresh (0 if none); one CTA per
// 64K elements would be overkill for a ladder rung: a single 1024-thread CTA strides over the vector (n <= 2^21)
__global__ void __launch_bounds__(1024) tail_kernel(const double* __restrict__ c, int64_t n, int64_t ntail, double thresh,
                                                    double* __restrict__ res) {
    AFB_DYN_SMEM(double, sm);   // 32 maxima, 32 tail maxima, 32 last indices
    double* smax = sm;
    double* stail = sm + 32;
    long long* slast = reinterpret_cast<long long*>(sm + 64);
    double mx = 0.0, tl = 0.0;
    long long last = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(c[i]);
        mx = fmax(mx, v);
        if (i >= n - ntail) tl = fmax(tl, v);
        if (v > thresh) last = i + 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        tl = fmax(tl, __shfl_xor_sync(0xffffffffu, tl, o));
        const long long ol = __shfl_xor_sync(0xffffffffu, last, o);
        last = ol > last ? ol : last;
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; stail[threadIdx.x >> 5] = tl; slast[threadIdx.x >> 5] = last; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            mx = fmax(mx, smax[w]); tl = fmax(tl, stail[w]); last = slast[w] > last ? slast[w] : last;
        }
        res[0] = mx; res[1] = tl; res[2] = (double)last;
    }
}

static inline unsigned fun_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = 8 * (int64_t)sm_count();
    return (unsigned)std::max<int64_t>(1, std::min(g, cap));
}

template <int OP> static int32_t launch_unary_op(const double* a, double* out, int64_t n, double alpha, double beta, void* stream) {
    auto k = unary_kernel<OP>;
    AFB_LAUNCH(k, fun_grid(n), 256, 0, stream, a, out, n, alpha, beta);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
template <int OP> static int32_t launch_binary_op(const double* a, const double* b, double* out, int64_t n, void* stream) {
    auto k = binary_kernel<OP>;
    AFB_LAUNCH(k, fun_grid(n), 256, 0, stream, a, b, out, n);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}

}  // namespace afb

using namespace afb;

struct BufGuard {   // makes the buffer's device current for the scope
    int keep = -1;
    explicit BufGuard(int dev) {
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) { keep = cur; cudaSetDevice(dev); }
    }
    ~BufGuard() { if (keep >= 0) cudaSetDevice(keep); }
};

static inline bool span_ok(afb_buf b, int64_t n, int64_t off = 0) { return b && n >= 0 && off >= 0 && off + n <= b->n; }

extern "C" {

int32_t afb_buf_create(afb_buf* out, int64_t n) {
    if (!out) return fail(AFB_BAD_ARG, "afb_buf_create: null out pointer");
    *out = nullptr;
    if (n < 0) return fail(AFB_BAD_SIZE, "afb_buf_create: negative size");
    AFB_TRY(ensure_device());
    afb_buf_s* b = new afb_buf_s();
    b->n = n;
    b->dev = current_device();
    // blocks are recycled through the per-device cache: every call of this layer runs on the legacy default stream, so a
    // block handed out again is only touched after the work of its previous owner
    void* d = nullptr;
    const int32_t st = device_block_get(sizeof(double) * (size_t)std::max<int64_t>(n, 2), &d, &b->cap);
    if (st != AFB_OK) { delete b; return st; }
    b->p = reinterpret_cast<double*>(d);
    *out = b;
    return AFB_OK;
}
int32_t afb_buf_destroy(afb_buf b) {
    if (!b) return AFB_OK;
    if (b->p) device_block_put(b->p, b->cap, b->dev);
    delete b;
    return AFB_OK;
}
int32_t afb_buf_size(afb_buf b, int64_t* n) {
    if (!b || !n) return fail(AFB_BAD_ARG, "afb_buf_size: null argument");
    *n = b->n;
    return AFB_OK;
}
int32_t afb_buf_ptr(afb_buf b, void** dptr) {
    if (!b || !dptr) return fail(AFB_BAD_ARG, "afb_buf_ptr: null argument");
    *dptr = b->p;
    return AFB_OK;
}
int32_t afb_buf_upload(afb_buf b, const double* host, int64_t n, int64_t offset) {
    if (!span_ok(b, n, offset)) return fail(AFB_BAD_SIZE, "afb_buf_upload: range outside the buffer");
    if (n == 0) return AFB_OK;
    if (!host) return fail(AFB_BAD_ARG, "afb_buf_upload: null host pointer");
    BufGuard g(b->dev);
    AFB_CUDA(cudaMemcpy(b->p + offset, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    return AFB_OK;
}
int32_t afb_buf_download(afb_buf b, double* host, int64_t n, int64_t offset) {
    if (!span_ok(b, n, offset)) return fail(AFB_BAD_SIZE, "afb_buf_download: range outside the buffer");
    if (n == 0) return AFB_OK;
    if (!host) return fail(AFB_BAD_ARG, "afb_buf_download: null host pointer");
    BufGuard g(b->dev);
    AFB_CUDA(cudaMemcpy(host, b->p + offset, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return AFB_OK;
}

int32_t afb_buf_unary(int32_t op, afb_buf a, afb_buf out, int64_t n, double alpha, double beta) {
    if (op < 0 || op >= UOP_COUNT) return fail(AFB_BAD_ARG, "afb_buf_unary: unknown op");
    if (!span_ok(a, n) || !span_ok(out, n)) return fail(AFB_BAD_SIZE, "afb_buf_unary: n exceeds a buffer");
    if (n == 0) return AFB_OK;
    if (a->dev != out->dev) return fail(AFB_BAD_ARG, "afb_buf_unary: buffers on different devices");
    BufGuard g(a->dev);
    switch (op) {
        case UOP_SIN: return launch_unary_op<UOP_SIN>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_COS: return launch_unary_op<UOP_COS>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_EXP: return launch_unary_op<UOP_EXP>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_ABS: return launch_unary_op<UOP_ABS>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_SQRT: return launch_unary_op<UOP_SQRT>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_SQUARE: return launch_unary_op<UOP_SQUARE>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_RECIP: return launch_unary_op<UOP_RECIP>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_LOG: return launch_unary_op<UOP_LOG>(a->p, out->p, n, alpha, beta, nullptr);
        case UOP_TANH: return launch_unary_op<UOP_TANH>(a->p, out->p, n, alpha, beta, nullptr);
        default: return launch_unary_op<UOP_IDENT>(a->p, out->p, n, alpha, beta, nullptr);
    }
}
int32_t afb_buf_binary(int32_t op, afb_buf a, afb_buf b, afb_buf out, int64_t n) {
    if (op < 0 || op >= BOP_COUNT) return fail(AFB_BAD_ARG, "afb_buf_binary: unknown op");
    if (!span_ok(a, n) || !span_ok(b, n) || !span_ok(out, n)) return fail(AFB_BAD_SIZE, "afb_buf_binary: n exceeds a buffer");
    if (n == 0) return AFB_OK;
    if (a->dev != out->dev || b->dev != out->dev) return fail(AFB_BAD_ARG, "afb_buf_binary: buffers on different devices");
    BufGuard g(a->dev);
    switch (op) {
        case BOP_ADD: return launch_binary_op<BOP_ADD>(a->p, b->p, out->p, n, nullptr);
        case BOP_SUB: return launch_binary_op<BOP_SUB>(a->p, b->p, out->p, n, nullptr);
        case BOP_MUL: return launch_binary_op<BOP_MUL>(a->p, b->p, out->p, n, nullptr);
        default: return launch_binary_op<BOP_DIV>(a->p, b->p, out->p, n, nullptr);
    }
}
int32_t afb_buf_pad(afb_buf in, int64_t n_in, afb_buf out, int64_t n_out) {
    if (!span_ok(in, n_in) || !span_ok(out, n_out)) return fail(AFB_BAD_SIZE, "afb_buf_pad: length exceeds a buffer");
    if (n_out == 0) return AFB_OK;
    if (in == out && n_out <= n_in) return AFB_OK;   // truncation in place is a no-op
    if (in->dev != out->dev) return fail(AFB_BAD_ARG, "afb_buf_pad: buffers on different devices");
    BufGuard g(in->dev);
    auto k = pad_kernel;
    AFB_LAUNCH(k, fun_grid(n_out), 256, 0, nullptr, in->p, n_in, out->p, n_out);
    AFB_KERNEL_CHECK();
    return AFB_OK;
}
int32_t afb_buf_chebpoints(afb_buf out, int64_t n, double a, double b) {
    if (!span_ok(out, n)) return fail(AFB_BAD_SIZE, "afb_buf_chebpoints: n exceeds the buffer");
    if (n == 0) return AFB_OK;
    BufGuard g(out->dev);
    return launch_chebpoints(out->p, n, a, b, nullptr);
}
int32_t afb_buf_clenshaw(afb_buf c, int64_t N, afb_buf x, afb_buf y, int64_t P) {
    if (!span_ok(c, N) || !span_ok(x, P) || !span_ok(y, P)) return fail(AFB_BAD_SIZE, "afb_buf_clenshaw: length exceeds a buffer");
    if (P == 0) return AFB_OK;
    BufGuard g(c->dev);
    return launch_clenshaw_vec(c->p, N, x->p, y->p, P, nullptr);
}
int32_t afb_buf_tail(afb_buf c, int64_t n, int64_t ntail, double rel_thresh, double* maxabs, double* tailmax, int64_t* chop_len) {
    if (!span_ok(c, n) || ntail < 0) return fail(AFB_BAD_SIZE, "afb_buf_tail: bad sizes");
    if (n == 0) { if (maxabs) *maxabs = 0; if (tailmax) *tailmax = 0; if (chop_len) *chop_len = 0; return AFB_OK; }
    BufGuard g(c->dev);
    void* d = nullptr;
    size_t dcap = 0;
    AFB_TRY(device_block_get(4 * sizeof(double), &d, &dcap));
    double h[3] = {0, 0, 0};
    auto k = tail_kernel;
    int32_t st = AFB_OK;
    // first the maxima (the chop threshold is relative to max |c|), then the chop length
    AFB_LAUNCH(k, 1, 1024, 96 * sizeof(double), nullptr, c->p, n, std::min(ntail, n), 1e300, reinterpret_cast<double*>(d));
    cudaError_t e = cudaMemcpy(h, d, 3 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) st = cuda_fail(e, "afb_buf_tail", __FILE__, __LINE__);
    if (st == AFB_OK && chop_len) {
        AFB_LAUNCH(k, 1, 1024, 96 * sizeof(double), nullptr, c->p, n, std::min(ntail, n), rel_thresh * h[0], reinterpret_cast<double*>(d));
        double h2[3];
        e = cudaMemcpy(h2, d, 3 * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = cuda_fail(e, "afb_buf_tail", __FILE__, __LINE__);
        else *chop_len = (int64_t)h2[2];
    }
    device_block_put(d, dcap, c->dev);
    if (st == AFB_OK) { if (maxabs) *maxabs = h[0]; if (tailmax) *tailmax = h[1]; }
    return st;
}
int32_t afb_plan_execute_buf(afb_plan plan, afb_buf in, afb_buf out) {
    if (!plan || !in || !out) return fail(AFB_BAD_ARG, "afb_plan_execute_buf: null argument");
    int64_t need_in = 0, need_out = 0;
    AFB_TRY(afb_plan_io_doubles(plan, &need_in, &need_out));
    if (in->n < need_in || out->n < need_out) return fail(AFB_BAD_SIZE, "afb_plan_execute_buf: buffer smaller than the plan's data");
    return afb_plan_execute_device(plan, in->p, out->p, nullptr);
}

}  // extern "C"
