/* Plain-C caller of libapproxfun_b200.so: what a non-Python host (Julia's ccall, a C program) does with the ABI of
 * include/approxfun_b200.h.  Exit codes: 0 = all checks passed, 3 = no CUDA device (the library has no CPU fallback),
 * 1 = a check failed.  Built and run by tests/test_abi.py (no GPU: expects 3) and tests/test_gpu_parity.py (expects 0). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/approxfun_b200.h"

static int die(const char* what, int32_t st) {
    char msg[512];
    afb_last_error(msg, sizeof msg);
    fprintf(stderr, "%s: status %d: %s\n", what, (int)st, msg);
    return st == AFB_NO_DEVICE ? 3 : 1;
}

int main(void) {
    enum { N = 64, P = 5 };
    double x[N], v[N], c[N], back[N], xs[P] = {-1.0, -0.3, 0.0, 0.7, 1.0}, ys[P];
    int32_t st = afb_init(-1);
    if (st != AFB_OK) return die("afb_init", st);
    printf("libapproxfun_b200 version %d\n", (int)afb_version());
    if ((st = afb_chebpoints(x, N, -1.0, 1.0)) != AFB_OK) return die("afb_chebpoints", st);
    for (int k = 0; k + 1 < N; ++k)
        if (!(x[k] > x[k + 1])) { fprintf(stderr, "points not descending\n"); return 1; }
    for (int k = 0; k < N; ++k) v[k] = 0.5 + 0.25 * x[k] + 2.0 * (4.0 * x[k] * x[k] * x[k] - 3.0 * x[k]); /* .5 T0 + .25 T1 + 2 T3 */
    afb_plan fwd = NULL, inv = NULL;
    if ((st = afb_plan_create(&fwd, AFB_CHEB1_FWD, N, 0, 1, N, 0)) != AFB_OK) return die("afb_plan_create", st);
    if ((st = afb_plan_execute_host(fwd, v, c)) != AFB_OK) return die("afb_plan_execute_host", st);
    for (int k = 0; k < N; ++k) {
        const double want = k == 0 ? 0.5 : k == 1 ? 0.25 : k == 3 ? 2.0 : 0.0;
        if (fabs(c[k] - want) > 1e-14) { fprintf(stderr, "coefficient %d = %.17g, want %g\n", k, c[k], want); return 1; }
    }
    if ((st = afb_plan_create(&inv, AFB_CHEB1_INV, N, 0, 1, N, 0)) != AFB_OK) return die("afb_plan_create(inv)", st);
    if ((st = afb_plan_execute_host(inv, c, back)) != AFB_OK) return die("afb_plan_execute_host(inv)", st);
    for (int k = 0; k < N; ++k)
        if (fabs(back[k] - v[k]) > 1e-13) { fprintf(stderr, "round trip failed at %d\n", k); return 1; }
    if ((st = afb_clenshaw_cheb(c, N, xs, ys, P)) != AFB_OK) return die("afb_clenshaw_cheb", st);
    for (int i = 0; i < P; ++i) {
        const double t = xs[i], want = 0.5 + 0.25 * t + 2.0 * (4.0 * t * t * t - 3.0 * t);
        if (fabs(ys[i] - want) > 1e-14) { fprintf(stderr, "clenshaw(%g) = %.17g, want %.17g\n", t, ys[i], want); return 1; }
    }
    afb_plan bad = NULL;
    if (afb_plan_create(&bad, AFB_PADUA_FWD, 5, 0, 1, 5, 0) != AFB_BAD_SIZE || bad != NULL) {
        fprintf(stderr, "Padua length 5 accepted\n");
        return 1;
    }
    afb_plan_destroy(fwd);
    afb_plan_destroy(inv);
    afb_shutdown();
    printf("abi smoke ok\n");
    return 0;
}
