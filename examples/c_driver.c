/* Plain-C caller of libamg1d's C ABI (include/amg1d.h) - no Python, no torch, no CUDA headers.
 *
 * Builds the two-level hierarchy of BASELINE config C1 in its simplest form by hand - the P1 finite
 * element Laplacian on n elements (Dirichlet at both ends, 1 x 1 "element blocks" = rows of a
 * tridiagonal matrix), pairwise aggregation as the transfer, the coarse operator by Galerkin product on
 * the GPU (amg1d_coarsen_level_galerkin) - and runs multigrid(H, x0, b, maxiter, tol) on it
 * (amg1d_solve, src/solvers.jl:116-139 of the reference).  A two-level aggregation method on the
 * Laplacian is no fast solver; the point here is the call sequence.
 *
 *   gcc -O2 -Iinclude examples/c_driver.c -Lagglomerationmultigrid1d_b200 -lamg1d \
 *       -Wl,-rpath,$PWD/agglomerationmultigrid1d_b200 -lm -o examples/c_driver && examples/c_driver 4096
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "amg1d.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != AMG1D_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, amg1d_last_error(h));          \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 4096;          /* fine unknowns (even) */
    const int maxiter = 200;
    if (n < 4 || n % 2) { fprintf(stderr, "n must be even and >= 4\n"); return 2; }
    amg1d_t* h = NULL;
    CHECK(amg1d_create(&h, 2, 0, NULL));

    double* lo = calloc(n, sizeof *lo), *di = calloc(n, sizeof *di), *up = calloc(n, sizeof *up);
    double* dinv = calloc(n, sizeof *dinv), *P = calloc(n, sizeof *P), *b = calloc(n, sizeof *b);
    double* x = calloc(n, sizeof *x), *res = calloc(maxiter, sizeof *res);
    int64_t* parent = calloc(n, sizeof *parent);
    if (!lo || !di || !up || !dinv || !P || !b || !x || !res || !parent) return 3;
    for (int64_t e = 0; e < n; ++e) {
        di[e] = 2.0;
        lo[e] = e > 0 ? -1.0 : 0.0;                               /* A_lo[0] = A_up[n-1] = 0 */
        up[e] = e < n - 1 ? -1.0 : 0.0;
        dinv[e] = 0.5;                                            /* JacobiSmoother: 1 / diag(A) */
        parent[e] = e / 2;
        P[e] = 1.0;
        b[e] = sin(3.14159265358979323846 * (double)(e + 1) / (double)(n + 1));
    }
    CHECK(amg1d_set_level(h, 0, n, 1, lo, di, up, dinv, 1, NULL, n));
    CHECK(amg1d_set_transfer(h, 0, n, 1, 1, parent, P, NULL));
    CHECK(amg1d_coarsen_level_galerkin(h, 0, n / 2, 1, NULL, 0)); /* level 1 = L' A L, point Jacobi */
    CHECK(amg1d_finalize(h));

    int iters = 0;
    CHECK(amg1d_solve(h, x, b, maxiter, 1e-8, 3, 3, 2.0 / 3.0, &iters, res, NULL, NULL));
    double bn = 0.0;
    for (int64_t e = 0; e < n; ++e) bn += b[e] * b[e];
    bn = sqrt(bn);
    printf("n = %lld: %d V-cycles, ||Ax-b|| / ||b|| = %.3e, kernel launches = %lld\n", (long long)n, iters,
           iters > 0 ? res[iters - 1] / bn : NAN, (long long)amg1d_get_info(h, "kernel_launches"));
    const int ok = iters > 0 && iters < maxiter && res[iters - 1] < 1e-8 * bn;   /* stop rule of multigrid() */
    CHECK(amg1d_destroy(h));
    free(lo); free(di); free(up); free(dinv); free(P); free(b); free(x); free(res); free(parent);
    puts(ok ? "OK" : "FAILED");
    return ok ? 0 : 1;
}
