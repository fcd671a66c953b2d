/* Plain-C caller of libapproxfun_b200.so: what a non-Python host (Julia's ccall, a C program) does with the ABI of
 * include/approxfun_b200.h.  Exit codes: 0 = all checks passed, 3 = no CUDA device (the library has no CPU fallback),
 * 1 = a check failed.  Built and run by tests/test_abi.py (no GPU: expects 3) and tests/test_gpu_parity.py (expects 0).
 * With two or more GPUs visible it also drives TWO of them from this one process (afb_init(devices, 2)): the batch of a 1-D
 * plan and a 2-D transform are sharded inside the library and must reproduce the one-device results. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/approxfun_b200.h"

static int die(const char* what, int32_t st) {
    char msg[512];
    afb_last_error(msg, sizeof msg);
    fprintf(stderr, "%s: status %d: %s\n", what, (int)st, msg);
    return st == AFB_NO_DEVICE ? 3 : 1;
}

/* one process, two GPUs: same calls, the library fans out */
static int two_gpus(void) {
    const int32_t one[1] = {0}, two[2] = {0, 1};
    int32_t st = afb_init(two, 2);
    if (st == AFB_BAD_ARG) { printf("one GPU visible: multi-device section skipped\n"); return 0; }
    if (st != AFB_OK) return die("afb_init(two devices)", st);
    enum { N1 = 1024, B1 = 2048, R2 = 1024, C2 = 2048 };
    double* v = (double*)malloc(sizeof(double) * N1 * B1);
    double* c1 = (double*)malloc(sizeof(double) * N1 * B1);
    double* c2 = (double*)malloc(sizeof(double) * N1 * B1);
    if (!v || !c1 || !c2) return 1;
    unsigned long long s = 88172645463325252ull;
    for (long i = 0; i < (long)N1 * B1; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    afb_plan p = NULL, q = NULL;
    int rc = 0;
    /* (1) batched 1-D: 2048 functions x 1024 (16 MiB) over two devices vs one device: bit-identical */
    if ((st = afb_plan_create(&p, AFB_CHEB1_FWD, N1, 0, B1, N1, 0)) != AFB_OK) return die("afb_plan_create", st);
    if ((st = afb_plan_execute_host(p, v, c2)) != AFB_OK) return die("afb_plan_execute_host (2 GPUs)", st);
    if ((st = afb_init(one, 1)) != AFB_OK) return die("afb_init(one device)", st);
    if ((st = afb_plan_execute_host(p, v, c1)) != AFB_OK) return die("afb_plan_execute_host (1 GPU)", st);
    for (long i = 0; i < (long)N1 * B1; ++i)
        if (c1[i] != c2[i]) { fprintf(stderr, "batch sharding over 2 GPUs differs at %ld\n", i); rc = 1; break; }
    /* (2) 2-D 1024 x 2048: slabs + peer-memory transpose over NVLink vs the one-device plan */
    if (!rc) {
        if ((st = afb_plan_create(&q, AFB_CHEB1_2D_FWD, R2, C2, 1, R2, 0)) != AFB_OK) return die("afb_plan_create(2D)", st);
        if ((st = afb_plan_execute_host(q, v, c1)) != AFB_OK) return die("2D (1 GPU)", st);
        if ((st = afb_init(two, 2)) != AFB_OK) return die("afb_init(two devices)", st);
        for (int rep = 0; rep < 3 && !rc; ++rep) {   /* repeated calls alternate the receive buffers */
            if ((st = afb_plan_execute_host(q, v, c2)) != AFB_OK) return die("2D (2 GPUs)", st);
            for (long i = 0; i < (long)R2 * C2; ++i)
                if (fabs(c1[i] - c2[i]) > 1e-13) { fprintf(stderr, "2-D over 2 GPUs differs at %ld: %.17g vs %.17g\n", i, c2[i], c1[i]); rc = 1; break; }
        }
    }
    afb_plan_destroy(p);
    afb_plan_destroy(q);
    free(v); free(c1); free(c2);
    if ((st = afb_init(one, 1)) != AFB_OK) return die("afb_init(one device)", st);
    if (!rc) printf("two GPUs from one process: batch sharding bit-identical, 2-D within 1e-13\n");
    return rc;
}

int main(void) {
    enum { N = 64, P = 5 };
    double x[N], v[N], c[N], back[N], xs[P] = {-1.0, -0.3, 0.0, 0.7, 1.0}, ys[P];
    int32_t st = afb_init(NULL, 0);
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
    {   /* a caller-owned array page-locked in place, and a length served by a split plan (8000 = 4 x 2000): round trip */
        enum { NS = 8000, NB = 600 };            /* 38 MB: above the 4 MiB threshold of the pageable ring */
        double* a = (double*)malloc(sizeof(double) * NS * NB);
        double* b = (double*)malloc(sizeof(double) * NS * NB);
        double* r = (double*)malloc(sizeof(double) * NS * NB);
        if (!a || !b || !r) return die("malloc", -1);
        for (long i = 0; i < (long)NS * NB; ++i) a[i] = sin(0.001 * (double)i) + 0.25 * cos(0.37 * (double)(i % 911));
        afb_plan f8 = NULL, i8 = NULL;
        if ((st = afb_plan_create(&f8, AFB_CHEB1_FWD, NS, 0, NB, NS, 0)) != AFB_OK) return die("afb_plan_create(8000)", st);
        if ((st = afb_plan_create(&i8, AFB_CHEB1_INV, NS, 0, NB, NS, 0)) != AFB_OK) return die("afb_plan_create(8000 inv)", st);
        if ((st = afb_plan_execute_host(f8, a, b)) != AFB_OK) return die("execute_host(8000, pageable)", st);
        if ((st = afb_host_register(a, sizeof(double) * NS * NB)) != AFB_OK) return die("afb_host_register", st);
        if ((st = afb_host_register(r, sizeof(double) * NS * NB)) != AFB_OK) return die("afb_host_register", st);
        if ((st = afb_plan_execute_host(f8, a, r)) != AFB_OK) return die("execute_host(8000, registered)", st);
        if (memcmp(b, r, sizeof(double) * NS * NB) != 0) { fprintf(stderr, "registered and pageable results differ\n"); return 1; }
        if ((st = afb_plan_execute_host(i8, b, r)) != AFB_OK) return die("execute_host(8000 inv)", st);
        if ((st = afb_host_unregister(a)) != AFB_OK || (st = afb_host_unregister(r)) != AFB_OK) return die("afb_host_unregister", st);
        for (long i = 0; i < (long)NS * NB; ++i)
            if (fabs(r[i] - a[i]) > 1e-12) { fprintf(stderr, "n = 8000 round trip failed at %ld\n", i); return 1; }
        afb_plan_destroy(f8);
        afb_plan_destroy(i8);
        free(a); free(b); free(r);
        printf("split plan n = 8000 and afb_host_register: ok\n");
    }
    {
        int rc = two_gpus();
        if (rc) return rc;
    }
    afb_shutdown();
    printf("abi smoke ok\n");
    return 0;
}
