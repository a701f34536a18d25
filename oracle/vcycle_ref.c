/* Oracle (TEST INFRASTRUCTURE, never shipped): plain-C restatement of the reference's hot path.
 *
 * Follows src/solvers.jl:19-50 (multigrid_v_cycle), :116-139 (the residual check of multigrid) and
 * src/smoother.jl:52-58, :69-81 (apply_smoother for JacobiSmoother / BlockJacobi) of
 * mheinz757/AgglomerationMultigrid1D, on the reference's own data contract: sparse level operators,
 * sparse interpolation matrices, one partial-pivoting LU per element block (getrf / getrs style),
 * freshly allocated temporaries for every expression as in the Julia code.
 *
 * Sparse products walk CSR rows: row i accumulates a_ij u_j in ascending j, which is exactly the
 * per-row summation order of the reference's CSC SpMV, so results equal the serial CSC product bit
 * for bit while rows can be distributed over OpenMP threads ("all host threads" arm).  L' r uses the
 * CSR form of L' (= CSC of L): one dot product per coarse DOF, as Julia's adjoint SpMV.
 * The coarsest level is solved with a dense partial-pivoting LU (it has 1-256 unknowns here;
 * the reference calls SuiteSparse).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n;                      /* DOFs */
    const int64_t *Ap, *Aj;         /* CSR of A */
    const double* Ax;
    int m;                          /* block size (0: point Jacobi) */
    int64_t nblocks;
    const int64_t* inds;            /* m x nblocks, column per block (mBlockInds) */
    double* lu;                     /* nblocks * m * m, row-major LU factors */
    int* piv;                       /* nblocks * m */
    const double* jac;              /* diagonal (point Jacobi) */
    /* transfer to the next coarser level */
    int64_t nc;
    const int64_t *Lp, *Lj, *LTp, *LTj;
    const double *Lx, *LTx;
} level_t;

typedef struct {
    int nlev;
    level_t* L;
    int64_t ncoarse;
    double* clu;                    /* dense LU of the coarsest operator */
    int* cpiv;
} ref_t;

static void csr_mv(int64_t n, const int64_t* p, const int64_t* j, const double* x, const double* u,
                   double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int64_t k = p[i]; k < p[i + 1]; ++k) s += x[k] * u[j[k]];
        y[i] = s;
    }
}

static int lu_factor(double* a, int* piv, int m) {
    for (int c = 0; c < m; ++c) {
        int p = c;
        double best = fabs(a[c * m + c]);
        for (int r = c + 1; r < m; ++r)
            if (fabs(a[r * m + c]) > best) { best = fabs(a[r * m + c]); p = r; }
        piv[c] = p;
        if (best == 0.0) return 1;
        if (p != c)
            for (int k = 0; k < m; ++k) { double t = a[c * m + k]; a[c * m + k] = a[p * m + k]; a[p * m + k] = t; }
        for (int r = c + 1; r < m; ++r) {
            a[r * m + c] /= a[c * m + c];
            const double f = a[r * m + c];
            for (int k = c + 1; k < m; ++k) a[r * m + k] -= f * a[c * m + k];
        }
    }
    return 0;
}

static void lu_solve(const double* a, const int* piv, int m, double* v) {
    for (int c = 0; c < m; ++c) { double t = v[c]; v[c] = v[piv[c]]; v[piv[c]] = t; }
    for (int r = 1; r < m; ++r) { double s = v[r]; for (int k = 0; k < r; ++k) s -= a[r * m + k] * v[k]; v[r] = s; }
    for (int r = m - 1; r >= 0; --r) {
        double s = v[r];
        for (int k = r + 1; k < m; ++k) s -= a[r * m + k] * v[k];
        v[r] = s / a[r * m + r];
    }
}

/* Y = alpha * S^-1 B  (fresh output vector, like the reference) */
static double* apply_smoother(const level_t* lv, const double* B, double alpha) {
    double* Y = (double*)calloc((size_t)lv->n, sizeof(double));
    if (lv->m == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < lv->n; ++i) Y[i] = alpha * (B[i] / lv->jac[i]);
        return Y;
    }
    const int m = lv->m;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < lv->nblocks; ++e) {
        double v[32];
        const int64_t* idx = lv->inds + e * m;
        for (int i = 0; i < m; ++i) v[i] = B[idx[i]];
        lu_solve(lv->lu + e * m * m, lv->piv + e * m, m, v);
        for (int i = 0; i < m; ++i) Y[idx[i]] += v[i];
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < lv->n; ++i) Y[i] = alpha * Y[i];
    return Y;
}

static double* residual(const level_t* lv, const double* rhs, const double* u) {
    double* r = (double*)malloc((size_t)lv->n * sizeof(double));
    csr_mv(lv->n, lv->Ap, lv->Aj, lv->Ax, u, r);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < lv->n; ++i) r[i] = rhs[i] - r[i];
    return r;
}

static void smooth(const level_t* lv, const double* rhs, double* u, int count, double alpha) {
    for (int s = 0; s < count; ++s) {
        double* r = residual(lv, rhs, u);
        double* y = apply_smoother(lv, r, alpha);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < lv->n; ++i) u[i] += y[i];
        free(r); free(y);
    }
}

int ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

ref_t* ref_create(int nlev) {
    ref_t* h = (ref_t*)calloc(1, sizeof(ref_t));
    h->nlev = nlev;
    h->L = (level_t*)calloc((size_t)nlev, sizeof(level_t));
    return h;
}

/* blocks: nblocks dense m x m diagonal blocks (row-major) to factorise; m = 0 -> jac diagonal */
int ref_set_level(ref_t* h, int l, int64_t n, const int64_t* Ap, const int64_t* Aj, const double* Ax,
                  int m, int64_t nblocks, const int64_t* inds, const double* blocks, const double* jac) {
    level_t* lv = &h->L[l];
    lv->n = n; lv->Ap = Ap; lv->Aj = Aj; lv->Ax = Ax; lv->m = m; lv->nblocks = nblocks; lv->inds = inds; lv->jac = jac;
    if (m > 32) return 2;
    if (m > 0) {
        lv->lu = (double*)malloc((size_t)nblocks * m * m * sizeof(double));
        lv->piv = (int*)malloc((size_t)nblocks * m * sizeof(int));
        memcpy(lv->lu, blocks, (size_t)nblocks * m * m * sizeof(double));
        for (int64_t e = 0; e < nblocks; ++e)
            if (lu_factor(lv->lu + e * m * m, lv->piv + e * m, m)) return 1;
    }
    return 0;
}

int ref_set_transfer(ref_t* h, int l, int64_t nc, const int64_t* Lp, const int64_t* Lj, const double* Lx,
                     const int64_t* LTp, const int64_t* LTj, const double* LTx) {
    level_t* lv = &h->L[l];
    lv->nc = nc; lv->Lp = Lp; lv->Lj = Lj; lv->Lx = Lx; lv->LTp = LTp; lv->LTj = LTj; lv->LTx = LTx;
    return 0;
}

int ref_set_coarse_dense(ref_t* h, int64_t n, const double* A) {
    h->ncoarse = n;
    h->clu = (double*)malloc((size_t)n * n * sizeof(double));
    h->cpiv = (int*)malloc((size_t)n * sizeof(int));
    memcpy(h->clu, A, (size_t)n * n * sizeof(double));
    return lu_factor(h->clu, h->cpiv, (int)n);
}

/* x: in x0, out x */
int ref_vcycle(ref_t* h, double* x, const double* b, int nPre, int nPost, double alpha) {
    const int n = h->nlev;
    double** u = (double**)calloc((size_t)n, sizeof(double*));
    double** rhs = (double**)calloc((size_t)n, sizeof(double*));
    u[0] = x;
    rhs[0] = (double*)b;
    for (int k = 0; k < n - 1; ++k) {
        level_t* lv = &h->L[k];
        if (k > 0) u[k] = (double*)calloc((size_t)lv->n, sizeof(double));
        smooth(lv, rhs[k], u[k], nPre, alpha);
        double* r = residual(lv, rhs[k], u[k]);
        rhs[k + 1] = (double*)malloc((size_t)lv->nc * sizeof(double));
        csr_mv(lv->nc, lv->LTp, lv->LTj, lv->LTx, r, rhs[k + 1]);
        free(r);
    }
    if (n > 1) u[n - 1] = (double*)malloc((size_t)h->ncoarse * sizeof(double));
    {
        double* c = (double*)malloc((size_t)h->ncoarse * sizeof(double));
        memcpy(c, rhs[n - 1], (size_t)h->ncoarse * sizeof(double));
        lu_solve(h->clu, h->cpiv, (int)h->ncoarse, c);
        memcpy(u[n - 1], c, (size_t)h->ncoarse * sizeof(double));
        free(c);
    }
    for (int k = n - 2; k >= 0; --k) {
        level_t* lv = &h->L[k];
        double* y = (double*)malloc((size_t)lv->n * sizeof(double));
        csr_mv(lv->n, lv->Lp, lv->Lj, lv->Lx, u[k + 1], y);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < lv->n; ++i) u[k][i] += y[i];
        free(y);
        smooth(lv, rhs[k], u[k], nPost, alpha);
    }
    for (int k = 1; k < n; ++k) { free(u[k]); free(rhs[k]); }
    free(u); free(rhs);
    return 0;
}

double ref_residual_norm(ref_t* h, const double* x, const double* b) {
    level_t* lv = &h->L[0];
    double* r = residual(lv, b, x);
    double s = 0.0;
    for (int64_t i = 0; i < lv->n; ++i) s += r[i] * r[i];
    free(r);
    return sqrt(s);
}

void ref_destroy(ref_t* h) {
    if (!h) return;
    for (int l = 0; l < h->nlev; ++l) { free(h->L[l].lu); free(h->L[l].piv); }
    free(h->L); free(h->clu); free(h->cpiv); free(h);
}
