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
 *
 * Block-pattern storage (ref_set_level_pattern / ref_set_transfer_pattern).  On the uniform meshes of the
 * BASELINE configs every level's sparse matrices repeat one interior block row; storing them as CSR at 2^24 -
 * 2^26 elements takes tens of GB.  The pattern mode holds the SAME sparse matrices as their n_head + 1 + n_tail
 * distinct block rows and walks each row exactly as the CSR code would: the non-zeros of a row in ascending
 * order of the reference's DOF number, explicit zeros skipped.  For DG-type levels (DOF (k-1)(p+1)+i,
 * src/dg_mesh.jl:41-46) that is block order; for CG levels (vertices numbered first, interior nodes after all
 * vertices, src/cg_mesh.jl:35-45) it is "all vertex columns of the row, then all interior-node columns"
 * (flag vertex_first), whatever grouping the vectors are held in.  tests/test_oracle_pattern.py checks the two
 * storage modes against each other bit for bit.  This is what lets the oracle judge the GPU at BASELINE's own
 * sizes (tests/test_gpu_atscale.py) and be timed there (bench.py --impl reference).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <malloc.h>
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
    /* block-pattern storage of A (pat = 1): ne element blocks of size pm, block row of element e =
       [lo | di | up][set(e)], each block pm x pm row-major; vfirst: CG numbering (see the header) */
    int pat, pm, pnh, pnt, vfirst;
    int64_t ne;
    const double *plo, *pdi, *pup;
    /* block-pattern storage of the transfer (tpat = 1): parent(e) = (e + shift) / ratio + base, blocks
       P0[blk(e)] (parent) and P1[blk(e)] (parent + 1, may be NULL), each tmf x tmc row-major */
    int tpat, tmf, tmc, ratio, shift, base, period, tnh, tnt, cvfirst;
    int64_t tnf, tnc;
    const double *P0, *P1;
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

/* ---- block-pattern storage: the same row walks as csr_mv -------------------------------------------- */
static inline int64_t pat_set(int64_t e, int64_t n, int nh, int nt) {
    return e < nh ? e : (e >= n - nt ? nh + 1 + (e - (n - nt)) : nh);
}

/* one pass over the columns [j0, j1) of the three blocks of row i, ascending block, skipping exact zeros */
#define PAT_ROW_PASS(j0, j1)                                                                      \
    do {                                                                                          \
        if (e > 0) for (int j = (j0); j < (j1); ++j) { const double a = lo[i * m + j]; if (a != 0.0) s += a * u[(e - 1) * m + j]; } \
        for (int j = (j0); j < (j1); ++j) { const double a = di[i * m + j]; if (a != 0.0) s += a * u[e * m + j]; }                  \
        if (e < ne - 1) for (int j = (j0); j < (j1); ++j) { const double a = up[i * m + j]; if (a != 0.0) s += a * u[(e + 1) * m + j]; } \
    } while (0)

static void pat_mv(const level_t* lv, const double* u, double* y) {
    const int m = lv->pm;
    const int64_t ne = lv->ne;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; ++e) {
        const int64_t st = pat_set(e, ne, lv->pnh, lv->pnt) * m * m;
        const double *lo = lv->plo + st, *di = lv->pdi + st, *up = lv->pup + st;
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            if (lv->vfirst) { PAT_ROW_PASS(0, 1); PAT_ROW_PASS(1, m); }   /* vertex columns, then interior ones */
            else PAT_ROW_PASS(0, m);
            y[e * m + i] = s;
        }
    }
}

static inline int64_t tr_blk(const level_t* lv, int64_t e) {
    if (e < lv->tnh) return e;
    if (e >= lv->tnf - lv->tnt) return lv->tnh + lv->period + (e - (lv->tnf - lv->tnt));
    return lv->tnh + (e - lv->tnh) % lv->period;
}
static inline int64_t tr_first(const level_t* lv, int64_t q) {   /* first fine element with parent >= q */
    int64_t v = (q - lv->base) * (int64_t)lv->ratio - lv->shift;
    return v < 0 ? 0 : (v > lv->tnf ? lv->tnf : v);
}

/* y = L uc: row (e, i) of L has its non-zeros in the blocks of parent(e) (P0) and parent(e) + 1 (P1) */
static void pat_prolong(const level_t* lv, const double* uc, double* y) {
    const int mf = lv->tmf, mc = lv->tmc;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < lv->tnf; ++e) {
        const int64_t par = (e + lv->shift) / lv->ratio + lv->base;
        const double* B0 = lv->P0 + tr_blk(lv, e) * mf * mc;
        const double* B1 = lv->P1 ? lv->P1 + tr_blk(lv, e) * mf * mc : NULL;
        const int ok0 = par >= 0 && par < lv->tnc, ok1 = B1 && par + 1 >= 0 && par + 1 < lv->tnc;
        for (int i = 0; i < mf; ++i) {
            double s = 0.0;
            const int npass = lv->cvfirst ? 2 : 1;
            for (int ps = 0; ps < npass; ++ps) {
                const int j0 = lv->cvfirst ? (ps == 0 ? 0 : 1) : 0, j1 = lv->cvfirst ? (ps == 0 ? 1 : mc) : mc;
                if (ok0) for (int j = j0; j < j1; ++j) { const double a = B0[i * mc + j]; if (a != 0.0) s += a * uc[par * mc + j]; }
                if (ok1) for (int j = j0; j < j1; ++j) { const double a = B1[i * mc + j]; if (a != 0.0) s += a * uc[(par + 1) * mc + j]; }
            }
            y[e * mf + i] = s;
        }
    }
}

/* rc = L' r: row (q, j) of L' collects the P1 blocks of the children of q - 1 and the P0 blocks of q's own */
static void pat_restrict(const level_t* lv, const double* r, double* rc) {
    const int mf = lv->tmf, mc = lv->tmc;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < lv->tnc; ++q) {
        const int64_t a0 = tr_first(lv, q - 1), a1 = tr_first(lv, q), a2 = tr_first(lv, q + 1);
        for (int j = 0; j < mc; ++j) {
            double s = 0.0;
            const int npass = lv->vfirst ? 2 : 1;
            for (int ps = 0; ps < npass; ++ps) {
                const int i0 = lv->vfirst ? (ps == 0 ? 0 : 1) : 0, i1 = lv->vfirst ? (ps == 0 ? 1 : mf) : mf;
                if (lv->P1)
                    for (int64_t e = a0; e < a1; ++e) {
                        const double* B = lv->P1 + tr_blk(lv, e) * mf * mc;
                        for (int i = i0; i < i1; ++i) { const double a = B[i * mc + j]; if (a != 0.0) s += a * r[e * mf + i]; }
                    }
                for (int64_t e = a1; e < a2; ++e) {
                    const double* B = lv->P0 + tr_blk(lv, e) * mf * mc;
                    for (int i = i0; i < i1; ++i) { const double a = B[i * mc + j]; if (a != 0.0) s += a * r[e * mf + i]; }
                }
            }
            rc[q * mc + j] = s;
        }
    }
}

static void op_mv(const level_t* lv, const double* u, double* y) {
    if (lv->pat) pat_mv(lv, u, y);
    else csr_mv(lv->n, lv->Ap, lv->Aj, lv->Ax, u, y);
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
    if (lv->pat) {          /* mBlockInds[:, e] = e m .. e m + m - 1; one LU per distinct block row */
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < lv->nblocks; ++e) {
            double v[32];
            const int64_t st = pat_set(e, lv->ne, lv->pnh, lv->pnt);
            for (int i = 0; i < m; ++i) v[i] = B[e * m + i];
            lu_solve(lv->lu + st * m * m, lv->piv + st * m, m, v);
            for (int i = 0; i < m; ++i) Y[e * m + i] += v[i];
        }
    } else {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < lv->nblocks; ++e) {
        double v[32];
        const int64_t* idx = lv->inds + e * m;
        for (int i = 0; i < m; ++i) v[i] = B[idx[i]];
        lu_solve(lv->lu + e * m * m, lv->piv + e * m, m, v);
        for (int i = 0; i < m; ++i) Y[idx[i]] += v[i];
    }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < lv->n; ++i) Y[i] = alpha * Y[i];
    return Y;
}

static double* residual(const level_t* lv, const double* rhs, const double* u) {
    double* r = (double*)malloc((size_t)lv->n * sizeof(double));
    op_mv(lv, u, r);
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
    /* every expression allocates a fresh vector, as in the Julia code; keep freed vectors in the heap
       instead of returning them to the kernel (munmap + page faults on every temporary would time the
       kernel's page allocator, which Julia's pooled GC heap does not pay either) */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, -1);
    mallopt(M_TOP_PAD, 64 << 20);
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

/* Block-pattern form of a level (see the header).  lo / di / up: (n_head + 1 + n_tail) blocks, m x m
 * row-major.  point_jacobi = 1: JacobiSmoother (diagonal of A); else BlockJacobi with one block per element. */
int ref_set_level_pattern(ref_t* h, int l, int64_t n_elem, int m, int n_head, int n_tail, const double* lo,
                          const double* di, const double* up, int point_jacobi, int vertex_first) {
    level_t* lv = &h->L[l];
    if (m < 1 || m > 32 || (int64_t)n_head + n_tail > n_elem) return 2;
    lv->pat = 1; lv->pm = m; lv->pnh = n_head; lv->pnt = n_tail; lv->ne = n_elem; lv->vfirst = vertex_first;
    lv->plo = lo; lv->pdi = di; lv->pup = up;
    lv->n = n_elem * m;
    const int ns = n_head + 1 + n_tail;
    if (point_jacobi) {
        lv->m = 0;
        double* jac = (double*)malloc((size_t)lv->n * sizeof(double));
        for (int64_t e = 0; e < n_elem; ++e) {
            const double* d = di + pat_set(e, n_elem, n_head, n_tail) * m * m;
            for (int i = 0; i < m; ++i) jac[e * m + i] = d[i * m + i];
        }
        lv->jac = jac;
        lv->lu = jac;          /* owned: released by ref_destroy */
    } else {
        lv->m = m; lv->nblocks = n_elem;
        lv->lu = (double*)malloc((size_t)ns * m * m * sizeof(double));
        lv->piv = (int*)malloc((size_t)ns * m * sizeof(int));
        memcpy(lv->lu, di, (size_t)ns * m * m * sizeof(double));
        for (int s = 0; s < ns; ++s)
            if (lu_factor(lv->lu + (size_t)s * m * m, lv->piv + (size_t)s * m, m)) return 1;
    }
    return 0;
}

/* Block-pattern form of the interpolation from level l + 1 to l.  coarse_vertex_first: numbering of level l + 1. */
int ref_set_transfer_pattern(ref_t* h, int l, int64_t n_fine, int64_t n_coarse, int mf, int mc, int ratio,
                             int shift, int base, int period, int n_head, int n_tail, const double* P0,
                             const double* P1, int coarse_vertex_first) {
    level_t* lv = &h->L[l];
    if (ratio < 1 || period < 1 || shift < 0 || shift >= ratio) return 2;
    lv->tpat = 1; lv->tnf = n_fine; lv->tnc = n_coarse; lv->tmf = mf; lv->tmc = mc; lv->ratio = ratio;
    lv->shift = shift; lv->base = base; lv->period = period; lv->tnh = n_head; lv->tnt = n_tail;
    lv->P0 = P0; lv->P1 = P1; lv->cvfirst = coarse_vertex_first;
    lv->nc = n_coarse * mc;
    return 0;
}

/* dense coarsest operator from the block rows of the last level (pattern mode) */
int ref_set_coarse_from_pattern(ref_t* h) {
    level_t* lv = &h->L[h->nlev - 1];
    if (!lv->pat) return 2;
    const int64_t n = lv->n;
    const int m = lv->pm;
    double* A = (double*)calloc((size_t)n * n, sizeof(double));
    for (int64_t e = 0; e < lv->ne; ++e) {
        const int64_t st = pat_set(e, lv->ne, lv->pnh, lv->pnt) * m * m;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                if (e > 0) A[(e * m + i) * n + (e - 1) * m + j] = lv->plo[st + i * m + j];
                A[(e * m + i) * n + e * m + j] = lv->pdi[st + i * m + j];
                if (e < lv->ne - 1) A[(e * m + i) * n + (e + 1) * m + j] = lv->pup[st + i * m + j];
            }
    }
    h->ncoarse = n;
    h->clu = A;
    h->cpiv = (int*)malloc((size_t)n * sizeof(int));
    return lu_factor(h->clu, h->cpiv, (int)n);
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
        if (lv->tpat) pat_restrict(lv, r, rhs[k + 1]);
        else csr_mv(lv->nc, lv->LTp, lv->LTj, lv->LTx, r, rhs[k + 1]);
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
        if (lv->tpat) pat_prolong(lv, u[k + 1], y);
        else csr_mv(lv->n, lv->Lp, lv->Lj, lv->Lx, u[k + 1], y);
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
    for (int l = 0; l < h->nlev; ++l) { free(h->L[l].lu); free(h->L[l].piv); }   /* (pattern point Jacobi: lu == jac) */
    free(h->L); free(h->clu); free(h->cpiv); free(h);
}
