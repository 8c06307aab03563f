/* abi_harness.c -- plain-C client of include/arrowhead_b200.h (test infrastructure).
 *
 *   abi_harness layout            prints sizeof / offsetof of every field of ah_result and ah_stats, one per line
 *   abi_harness solve <lib.so>    dlopen()s the library exactly as Julia's ccall does, solves one small SymArrow, SymDPR1 and
 *                                 HalfArrow through the C ABI and prints every output as hex floats / integers
 *
 * tests/test_abi_layout.py compares `layout` with the ctypes mirror and with the Julia struct of ArrowheadB200.jl (CPU),
 * and `solve` with the same calls made through ctypes (GPU): the header is usable from C as written, and the three
 * bindings agree on what the bytes mean. */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arrowhead_b200.h"

#define OFF(T, f) printf(#T "." #f " %zu %zu\n", offsetof(T, f), sizeof(((T *)0)->f))

static int layout(void) {
    printf("sizeof.ah_result %zu\n", sizeof(ah_result));
    OFF(ah_result, lambda); OFF(ah_result, sind); OFF(ah_result, kb); OFF(ah_result, kz); OFF(ah_result, knu);
    OFF(ah_result, krho); OFF(ah_result, qout); OFF(ah_result, steps); OFF(ah_result, status);
    OFF(ah_result, U); OFF(ah_result, ldu); OFF(ah_result, U_mem);
    OFF(ah_result, V); OFF(ah_result, ldv); OFF(ah_result, V_mem);
    OFF(ah_result, rowmap_u); OFF(ah_result, rowmap_v); OFF(ah_result, rowscale_v);
    printf("sizeof.ah_stats %zu\n", sizeof(ah_stats));
    OFF(ah_stats, kernel_ms); OFF(ah_stats, main_kernel_ms); OFF(ah_stats, h2d_ms); OFF(ah_stats, d2h_ms);
    OFF(ah_stats, launches); OFF(ah_stats, n_fast); OFF(ah_stats, n_full); OFF(ah_stats, total_steps);
    OFF(ah_stats, h2d_bytes); OFF(ah_stats, d2h_bytes);
    OFF(ah_stats, prep_ms); OFF(ah_stats, bisect_ms); OFF(ah_stats, vec_ms); OFF(ah_stats, full_ms);
    OFF(ah_stats, n_rem3); OFF(ah_stats, rem3_ms); OFF(ah_stats, host_pre_ms); OFF(ah_stats, host_post_ms);
    return 0;
}

typedef int (*create_fn)(ah_ctx **, int);
typedef void (*destroy_fn)(ah_ctx *);
typedef const char *(*err_fn)(const ah_ctx *);
typedef int (*arrow_fn)(ah_ctx *, int64_t, const double *, const double *, double, const double *, int64_t, int64_t, ah_result *);
typedef int (*half_fn)(ah_ctx *, int64_t, int64_t, const double *, const double *, const double *, int64_t, int64_t, ah_result *);

static void dump(const char *tag, int64_t cnt, const ah_result *r, int64_t nu, int64_t nv) {
    for (int64_t c = 0; c < cnt; ++c)
        printf("%s k=%lld lambda=%a sind=%lld kb=%a kz=%a knu=%a krho=%a qout=%lld status=%d\n", tag, (long long)(c + 1), r->lambda[c],
               (long long)r->sind[c], r->kb[c], r->kz[c], r->knu[c], r->krho[c], (long long)r->qout[c], r->status[c]);
    for (int64_t c = 0; c < cnt; ++c) {
        printf("%s U%lld", tag, (long long)(c + 1));
        for (int64_t l = 0; l < nu; ++l) printf(" %a", r->U[c * r->ldu + l]);
        printf("\n");
        if (r->V) {
            printf("%s V%lld", tag, (long long)(c + 1));
            for (int64_t l = 0; l < nv; ++l) printf(" %a", r->V[c * r->ldv + l]);
            printf("\n");
        }
    }
}

static ah_result make(int64_t cnt, int64_t nu, int64_t nv) {
    ah_result r;
    memset(&r, 0, sizeof r);
    r.lambda = calloc(cnt, 8); r.sind = calloc(cnt, 8); r.kb = calloc(cnt, 8); r.kz = calloc(cnt, 8); r.knu = calloc(cnt, 8);
    r.krho = calloc(cnt, 8); r.qout = calloc(cnt, 8); r.steps = calloc(cnt, 4); r.status = calloc(cnt, 4);
    r.U = calloc(cnt * nu, 8); r.ldu = nu; r.U_mem = AH_MEM_HOST;
    if (nv) { r.V = calloc(cnt * nv, 8); r.ldv = nv; r.V_mem = AH_MEM_HOST; }
    return r;
}

static int solve(const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    create_fn create = (create_fn)dlsym(h, "ah_create");
    destroy_fn destroy = (destroy_fn)dlsym(h, "ah_destroy");
    err_fn lasterr = (err_fn)dlsym(h, "ah_last_error");
    arrow_fn arrow = (arrow_fn)dlsym(h, "ah_symarrow_eigen"), dpr1 = (arrow_fn)dlsym(h, "ah_symdpr1_eigen");
    half_fn half = (half_fn)dlsym(h, "ah_halfarrow_svd");
    if (!create || !destroy || !lasterr || !arrow || !dpr1 || !half) { fprintf(stderr, "missing symbol\n"); return 2; }
    ah_ctx *ctx = NULL;
    int rc = create(&ctx, 0);
    if (rc) { fprintf(stderr, "ah_create rc=%d: %s\n", rc, lasterr(NULL)); return 3; }
    /* runtests.jl:22-25 Example 1 in its ordered form */
    const double D[5] = {2e-3, 1e-7, 0.0, -1e-7, -2e-3}, z[5] = {1e7, 1e7, 1.0, 1e7, 1e7};
    ah_result r = make(6, 6, 0);
    rc = arrow(ctx, 5, D, z, 1e20, NULL, 1, 6, &r);
    if (rc) { fprintf(stderr, "ah_symarrow_eigen rc=%d: %s\n", rc, lasterr(ctx)); return 4; }
    dump("symarrow", 6, &r, 6, 0);
    const double D2[4] = {10.0, 5.0, 1.0, -2.0}, u2[4] = {1.0, 0.5, -0.25, 2.0};
    ah_result r2 = make(4, 4, 0);
    rc = dpr1(ctx, 4, D2, u2, 1.5, NULL, 1, 4, &r2);
    if (rc) { fprintf(stderr, "ah_symdpr1_eigen rc=%d: %s\n", rc, lasterr(ctx)); return 4; }
    dump("symdpr1", 4, &r2, 4, 0);
    const double D3[3] = {3.0, 2.0, 0.5}, z3[4] = {1.0, -0.5, 0.25, 2.0};
    ah_result r3 = make(4, 4, 4);
    rc = half(ctx, 3, 4, D3, z3, NULL, 1, 4, &r3);
    if (rc) { fprintf(stderr, "ah_halfarrow_svd rc=%d: %s\n", rc, lasterr(ctx)); return 4; }
    dump("halfarrow", 4, &r3, 4, 4);
    /* contract violations come back as codes, nothing aborts */
    const double Dbad[2] = {1.0, 2.0};
    rc = arrow(ctx, 2, Dbad, z, 0.0, NULL, 1, 3, &r);
    printf("unsorted rc=%d\n", rc);
    destroy(ctx);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
    if (argc >= 3 && !strcmp(argv[1], "solve")) return solve(argv[2]);
    fprintf(stderr, "usage: %s layout | solve <libarrowhead_b200.so>\n", argv[0]);
    return 1;
}
