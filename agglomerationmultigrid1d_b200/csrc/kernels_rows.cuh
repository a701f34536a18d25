// Fused multi-sweep legs for LARGE element blocks (m = 5 .. 9: the DG p = 8 / CG p = 8 levels the
// reference's own hierarchy scripts start from, tests/dg_heirarchy_test.jl, tests/dg_cg_heirarchy_test.jl).
//
// kernels_fused.cuh keeps a whole element (its A_lo / A_di / A_up blocks) in ONE thread's registers,
// which stops at m = 5 (a dense 9 x 9 block alone is 81 doubles).  Here an element is spread over M
// threads, one per block ROW:
//
//   * a CTA owns a window of W consecutive elements (W = 32 or 64: one or two element tiles) and has
//     M * W threads; thread t = i * W + w owns row i of window element w, so a warp is "row i of one
//     element tile" and every operator load is the same fully used 256-byte request as in the
//     one-thread-per-element kernels (layout.cuh);
//   * the thread keeps row i of A_di (M doubles) and its part of A_lo / A_up in registers for the
//     whole leg; row i of Dinv (M doubles, used once per sweep) and the single trace row of the
//     compressed structure classes are staged global -> shared with cp.async into thread-private
//     columns (no barrier needed for them);
//   * per sweep the M threads of an element exchange the iterate through shared memory (double
//     buffered: Jacobi needs the old values), and for block smoothers also the residual (z = Dinv r
//     needs all of r);
//   * vectors (b, x in, x out) move through a coalesced window copy: thread t handles the t-th double
//     of the window's contiguous range, the shared-memory stage does the (element, row) transpose;
//     the prolongation of f_up is applied on that coalesced mapping.
//
// Row i of every product is accumulated in exactly the order of g_row_Ax / g_row_Dinv / g_restrict /
// g_prolong (kernels_generic.cuh), so the iterates are bit-identical to the generic tier - the same
// guarantee the one-thread-per-element kernels give.  Window / halo logic, the 32-bit index constants
// (WinIdx), the slab arguments of the sharded case and programmatic dependent launch are those of
// f_down / f_up.
//
// Algorithmic bytes per element: 8 (Kop + 3 M + MC / R), the operator is read ONCE per leg (the
// streaming path reads it nPre + 1 times on the way down and nPost (+ 1) times on the way up).
#pragma once
#include <cuda_runtime.h>

#include "kernels_fused.cuh"

// Resident threads per SM the launch bounds aim for: 1152 (<= 56 registers per thread) for the
// compressed structure classes, 768 (<= 80 registers) for dense off-diagonal blocks, whose rows of
// A_lo and A_up are register resident too (3 M doubles).
#ifndef ROWS_THREADS_PER_SM
#define ROWS_THREADS_PER_SM 1152
#endif
constexpr int rows_min_blocks(int m, int w, int st) {
    const int target = st == AMG1D_ST_DENSE ? 768 : ROWS_THREADS_PER_SM;
    return target / (m * w) >= 1 ? target / (m * w) : 1;
}

template <int M, int W, int ST, bool DIAG>
struct RowsSmem {
    static constexpr int XS = W + 3;          // slots 0 and W + 1 are the (zero) window edges
    static constexpr int RS = W + 9;
    double xs[2][M][XS];                      // iterate exchange, xs[buf][row][window position + 1]
    double rs[M][RS];                         // b stage, residual exchange, restriction source
    double dv[DIAG ? 1 : M][M * W];           // dv[j][t] = Dinv[i, j] of thread t's element (block smoother)
    double sv[ST == AMG1D_ST_DENSE ? 1 : M][W];  // the one stored row of A_up (ST_COLROW) / A_lo (ST_ROWCOL)
};

// what thread (row i) keeps in registers
template <int M, int ST>
struct RowOp {
    double di[M];                                     // A_di[i, :]
    double lo[ST == AMG1D_ST_DENSE ? M : 1];          // A_lo[i, :]  |  ST_COLROW: A_lo[i, ilo]  |  ST_ROWCOL: unused
    double up[ST == AMG1D_ST_DENSE ? M : 1];          // A_up[i, :]  |  ST_ROWCOL: A_up[i, iup]  |  ST_COLROW: unused
    double dinv;                                      // point Jacobi: Dinv[i]
};

template <int M, int W, int ST, bool DIAG>
__device__ __forceinline__ void rows_load(const double* __restrict__ mat, int64_t e, bool active, int ilo,
                                          int iup, int i, int w, int t, RowOp<M, ST>& A,
                                          RowsSmem<M, W, ST, DIAG>& S) {
    using O = OpShape<M, ST>;
    constexpr int ND = DIAG ? M : M * M;
    constexpr int K = O::O_DV + ND;
    A.dinv = 0.0;
    if (active) {
        const double* T = mat + (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31);
        if constexpr (DIAG) {
            A.dinv = T[(O::O_DV + i) * AMG1D_TILE];
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j) cp_async8(&S.dv[j][t], T + (O::O_DV + j * M + i) * AMG1D_TILE);
        }
#pragma unroll
        for (int j = 0; j < M; ++j) A.di[j] = T[(O::O_DI + j * M + i) * AMG1D_TILE];
        if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
            for (int j = 0; j < M; ++j) A.lo[j] = T[(j * M + i) * AMG1D_TILE];
#pragma unroll
            for (int j = 0; j < M; ++j) A.up[j] = T[(O::O_UP + j * M + i) * AMG1D_TILE];
        } else if constexpr (ST == AMG1D_ST_COLROW) {
            A.lo[0] = T[i * AMG1D_TILE];
            A.up[0] = 0.0;
            if (i == iup) {
#pragma unroll
                for (int j = 0; j < M; ++j) cp_async8(&S.sv[j][w], T + (O::O_UP + j) * AMG1D_TILE);
            }
        } else {
            A.up[0] = T[(O::O_UP + i) * AMG1D_TILE];
            A.lo[0] = 0.0;
            if (i == ilo) {
#pragma unroll
                for (int j = 0; j < M; ++j) cp_async8(&S.sv[j][w], T + j * AMG1D_TILE);
            }
        }
    } else {
        if constexpr (!DIAG) {
#pragma unroll
            for (int j = 0; j < M; ++j) S.dv[j][t] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < M; ++j) A.di[j] = 0.0;
        if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
            for (int j = 0; j < M; ++j) { A.lo[j] = 0.0; A.up[j] = 0.0; }
        } else {
            A.lo[0] = 0.0;
            A.up[0] = 0.0;
            if (i == (ST == AMG1D_ST_COLROW ? iup : ilo)) {
#pragma unroll
                for (int j = 0; j < M; ++j) S.sv[j][w] = 0.0;
            }
        }
    }
}

// row i of  A_lo x[w - 1] + A_di x[w] + A_up x[w + 1]  from exchange buffer `buf`; order of g_row_Ax
template <int M, int W, int ST, bool DIAG>
__device__ __forceinline__ double rows_Ax(const RowOp<M, ST>& A, const RowsSmem<M, W, ST, DIAG>& S, int buf,
                                          int ilo, int iup, int i, int w) {
    double y = 0.0;
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j) y = fma(A.lo[j], S.xs[buf][j][w], y);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
        y = fma(A.lo[0], S.xs[buf][ilo][w], 0.0);
    } else {
        if (i == ilo) {
#pragma unroll
            for (int j = 0; j < M; ++j) y = fma(S.sv[j][w], S.xs[buf][j][w], y);
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) y = fma(A.di[j], S.xs[buf][j][w + 1], y);
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j) y = fma(A.up[j], S.xs[buf][j][w + 2], y);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
        if (i == iup) {
#pragma unroll
            for (int j = 0; j < M; ++j) y = fma(S.sv[j][w], S.xs[buf][j][w + 2], y);
        }
    } else {
        y = fma(A.up[0], S.xs[buf][iup][w + 2], y);
    }
    return y;
}

// x_i += alpha (Dinv r)_i; block smoothers exchange r among the rows of the element (one barrier)
template <int M, int W, int ST, bool DIAG>
__device__ __forceinline__ double rows_update(const RowOp<M, ST>& A, RowsSmem<M, W, ST, DIAG>& S, double r,
                                              double xc, double alpha, int i, int w, int t) {
    if constexpr (DIAG) {
        return __dadd_rn(xc, __dmul_rn(alpha, A.dinv * r));
    } else {
        S.rs[i][w] = r;
        __syncthreads();
        double z = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) z = fma(S.dv[j][t], S.rs[j][w], z);
        return __dadd_rn(xc, __dmul_rn(alpha, z));
    }
}

template <int M, int W, int ST, bool DIAG>
__device__ __forceinline__ void rows_init_edges(RowsSmem<M, W, ST, DIAG>& S) {
    if (threadIdx.x < 4 * M) {
        const int buf = threadIdx.x / (2 * M);
        const int i = (threadIdx.x >> 1) % M;
        S.xs[buf][i][(threadIdx.x & 1) ? W + 1 : 0] = 0.0;
    }
}

#define ROWS_BOUNDS __launch_bounds__(M * W, rows_min_blocks(M, W, ST))

// nsweep pre-smoothing sweeps + residual + restriction (f_down for large blocks)
template <int M, int MC, int W, int ST, bool DIAG>
__global__ void ROWS_BOUNDS
r_down(const double* __restrict__ mat, int ilo, int iup, const double* __restrict__ b,
       const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
       const double* __restrict__ P1, TransferMap tm, double* __restrict__ rc, int64_t n, double alpha,
       int nsweep, int zero_guess, WinIdx wi, Slab sl) {
    extern __shared__ __align__(16) double rows_smem[];
    RowsSmem<M, W, ST, DIAG>& S = *reinterpret_cast<RowsSmem<M, W, ST, DIAG>*>(rows_smem);
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int w = t % W, i = t / W;
    const int halo = wi.halo, out = wi.out;
    const int64_t e0 = (int64_t)blockIdx.x * out - halo;          // local index of window element 0
    const int64_t e = e0 + w;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    rows_init_edges<M, W, ST, DIAG>(S);
    RowOp<M, ST> A;
    rows_load<M, W, ST, DIAG>(mat, e, active, ilo, iup, i, w, t, A, S);   // independent of earlier kernels
    pdl_wait();
    const int we = t / M, ie = t - we * M;                        // coalesced mapping: t-th double of the window
    const int64_t ee = e0 + we;
    {
        const bool act = ee >= -(int64_t)sl.gl && ee < n + sl.gr;
        S.rs[ie][we] = act ? b[e0 * M + t] : 0.0;
        S.xs[0][ie][we + 1] = (act && !zero_guess) ? xin[e0 * M + t] : 0.0;
    }
    cp_async_commit_wait_all();
    __syncthreads();
    const double bb = S.rs[i][w];
    double xc = S.xs[0][i][w + 1];
    int buf = 0;
    for (int s = 0; s < nsweep; ++s) {
        double r;
        if (zero_guess && s == 0) r = bb - 0.0;
        else r = bb - rows_Ax<M, W, ST, DIAG>(A, S, buf, ilo, iup, i, w);
        xc = rows_update<M, W, ST, DIAG>(A, S, r, xc, alpha, i, w, t);
        S.xs[buf ^ 1][i][w + 1] = xc;
        __syncthreads();
        buf ^= 1;
    }
    if (we >= halo && we < halo + out && ee < n) xout[e0 * M + t] = S.xs[buf][ie][we + 1];
    // residual with the final iterate, then restriction (order of g_restrict)
    {
        const double r = bb - rows_Ax<M, W, ST, DIAG>(A, S, buf, ilo, iup, i, w);
        S.rs[i][w] = r;
    }
    __syncthreads();
    const bool mine = w >= halo && w < halo + out;
    const int64_t eg = e + sl.e_off;
    if (i < MC && mine && eg < tm.n_fine) {
        int kdiv, kmod;
        small_divmod(wi.qmod0 + w, tm.ratio, &kdiv, &kmod);
        if (kmod == 0) {
            const int64_t Kc = (int64_t)blockIdx.x * wi.opr + wi.qdiv0 + kdiv + tm.base;   // eg == tm.first(Kc)
            const int64_t Kl = Kc - sl.c_off;
            if (Kc >= 0 && Kc < tm.n_coarse && Kl >= 0 && Kl < sl.nc) {
                double acc = 0.0;
                if (P1) {                                       // children of Kc - 1: [eg - ratio, eg), clamped
                    const int k0 = eg >= tm.ratio ? -tm.ratio : -(int)eg;
                    for (int k = k0; k < 0; ++k) {
                        const double* P = P1 + win_blk(tm, wi, eg + k, w + k) * (M * MC) + i * M;
#pragma unroll
                        for (int q = 0; q < M; ++q) acc = fma(P[q], S.rs[q][w + k], acc);
                    }
                }
                const int k1 = eg + tm.ratio <= tm.n_fine ? tm.ratio : (int)(tm.n_fine - eg);   // own children
                for (int k = 0; k < k1; ++k) {
                    const double* P = P0 + win_blk(tm, wi, eg + k, w + k) * (M * MC) + i * M;
#pragma unroll
                    for (int q = 0; q < M; ++q) acc = fma(P[q], S.rs[q][w + k], acc);
                }
                rc[Kl * MC + i] = acc;
            }
        }
    }
}

// prolongation + correction + nsweep post-smoothing sweeps (+ || b - A x ||^2 partial sums)
template <int M, int MC, int W, int ST, bool DIAG>
__global__ void ROWS_BOUNDS
r_up(const double* __restrict__ mat, int ilo, int iup, const double* __restrict__ b,
     const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
     const double* __restrict__ P1, TransferMap tm, const double* __restrict__ xcoarse, int64_t n,
     double alpha, int nsweep, WinIdx wi, double* __restrict__ partial, Slab sl) {
    extern __shared__ __align__(16) double rows_smem[];
    RowsSmem<M, W, ST, DIAG>& S = *reinterpret_cast<RowsSmem<M, W, ST, DIAG>*>(rows_smem);
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int w = t % W, i = t / W;
    const int halo = wi.halo, out = wi.out;
    const int64_t e0 = (int64_t)blockIdx.x * out - halo;
    const int64_t e = e0 + w;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    rows_init_edges<M, W, ST, DIAG>(S);
    RowOp<M, ST> A;
    rows_load<M, W, ST, DIAG>(mat, e, active, ilo, iup, i, w, t, A, S);
    pdl_wait();
    const int we = t / M, ie = t - we * M;
    const int64_t ee = e0 + we;
    {
        double bv = 0.0, xv = 0.0;
        if (ee >= -(int64_t)sl.gl && ee < n + sl.gr) {
            bv = b[e0 * M + t];
            xv = xin[e0 * M + t];
            // x += P0 x_c[parent] (+ P1 x_c[parent + 1]), row ie of element ee (order of g_prolong)
            const int64_t pb = win_blk(tm, wi, ee + sl.e_off, we) * (M * MC);
            int kdiv, kmod;
            small_divmod(wi.qmod0 + we, tm.ratio, &kdiv, &kmod);
            const int64_t par = (int64_t)blockIdx.x * wi.opr + wi.qdiv0 + kdiv + tm.base;
            const double* c0 = xcoarse + (par - sl.c_off) * MC;
            double y = 0.0;
#pragma unroll
            for (int j = 0; j < MC; ++j) y = fma(P0[pb + j * M + ie], c0[j], y);
            if (P1) {
#pragma unroll
                for (int j = 0; j < MC; ++j) y = fma(P1[pb + j * M + ie], c0[MC + j], y);
            }
            xv = xv + y;
        }
        S.rs[ie][we] = bv;
        S.xs[0][ie][we + 1] = xv;
    }
    cp_async_commit_wait_all();
    __syncthreads();
    const double bb = S.rs[i][w];
    double xc = S.xs[0][i][w + 1];
    int buf = 0;
    for (int s = 0; s < nsweep; ++s) {
        const double r = bb - rows_Ax<M, W, ST, DIAG>(A, S, buf, ilo, iup, i, w);
        xc = rows_update<M, W, ST, DIAG>(A, S, r, xc, alpha, i, w, t);
        S.xs[buf ^ 1][i][w + 1] = xc;
        __syncthreads();
        buf ^= 1;
    }
    if (we >= halo && we < halo + out && ee < n) xout[e0 * M + t] = S.xs[buf][ie][we + 1];
    if (partial) {
        const double r = bb - rows_Ax<M, W, ST, DIAG>(A, S, buf, ilo, iup, i, w);
        S.rs[i][w] = r;
        __syncthreads();
        double s2 = 0.0;
        if (i == 0 && e < n && w >= halo && w < halo + out) {
#pragma unroll
            for (int q = 0; q < M; ++q) s2 = fma(S.rs[q][w], S.rs[q][w], s2);
        }
        s2 = block_sum(s2);
        if (t == 0) partial[blockIdx.x] = s2;
    }
}

// ---- host-side dispatch ---------------------------------------------------------------------------------
// (M, MC, ST, DIAG): block Jacobi on DG levels p = 5 .. 8 (-> div(p, 2), or straight to pAgg = 1) and
// p = 4 -> pAgg = 1, dense and DG-assembled structure; point Jacobi on CG levels p = 5 .. 8 in group form.
#define ROWS_COMBOS(X)                                                                               \
    X(9, 5, 0, false) X(9, 5, 1, false) X(9, 2, 0, false) X(9, 2, 1, false)                          \
    X(8, 4, 0, false) X(8, 4, 1, false) X(7, 4, 0, false) X(7, 4, 1, false)                          \
    X(6, 3, 0, false) X(6, 3, 1, false) X(5, 2, 0, false) X(5, 2, 1, false)                          \
    X(8, 4, 2, true) X(7, 3, 2, true) X(6, 3, 2, true) X(5, 2, 2, true)

inline bool rows_window_ok(int window) { return window == 32 || window == 64; }

template <int M, int MC, int W, int ST, bool DIAG>
inline cudaError_t rows_configure_t() {
    const int smem = (int)sizeof(RowsSmem<M, W, ST, DIAG>);
    cudaError_t e = cudaFuncSetAttribute(r_down<M, MC, W, ST, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(r_up<M, MC, W, ST, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// once per device context and level shape, before the first launch (amg1d_finalize); *have = a kernel exists
inline cudaError_t rows_configure(const MatDesc& d, int mc, bool* have) {
    *have = true;
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0): {                                                \
        cudaError_t e = rows_configure_t<MM, MCC, 32, SS, DG>();                                         \
        return e != cudaSuccess ? e : rows_configure_t<MM, MCC, 64, SS, DG>();                           \
    }
        ROWS_COMBOS(X)
#undef X
        default: *have = false; return cudaSuccess;
    }
}

// Same contract as fused_down / fused_up (FUSED_NA / FUSED_OK / FUSED_ERR).
inline int rows_down(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, bool zero,
                     const double* mat, const double* b, const double* xin, double* xout,
                     const double* P0, const double* P1, double* rc, int64_t n, int64_t n_cover,
                     double alpha, const Slab& sl, int window, cudaStream_t st, bool pdl, cudaError_t* err) {
    if (!rows_window_ok(window)) return FUSED_NA;
    const WinIdx w = fused_window(nsweep, tm, P1 != nullptr || tm.shift != 0 || tm.base != 0, sl, window);
    if (w.out < tm.ratio || w.out < window / 2) return FUSED_NA;
    const unsigned grid = (unsigned)((n_cover + w.out - 1) / w.out);
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0):                                                  \
        if (window == 32)                                                                                \
            *err = launch_fused(r_down<MM, MCC, 32, SS, DG>, grid, MM * 32, sizeof(RowsSmem<MM, 32, SS, DG>), st, \
                                pdl, mat, d.ilo, d.iup, b, xin, xout, P0, P1, tm, rc, n, alpha, nsweep,  \
                                zero ? 1 : 0, w, sl);                                                    \
        else                                                                                             \
            *err = launch_fused(r_down<MM, MCC, 64, SS, DG>, grid, MM * 64, sizeof(RowsSmem<MM, 64, SS, DG>), st, \
                                pdl, mat, d.ilo, d.iup, b, xin, xout, P0, P1, tm, rc, n, alpha, nsweep,  \
                                zero ? 1 : 0, w, sl);                                                    \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;
        ROWS_COMBOS(X)
#undef X
        default: return FUSED_NA;
    }
}

inline int rows_up(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, const double* mat,
                   const double* b, const double* xin, double* xout, const double* P0, const double* P1,
                   const double* xcoarse, int64_t n, double alpha, double* partial, int64_t partial_cap,
                   int* nblocks, const Slab& sl, int window, cudaStream_t st, bool pdl, cudaError_t* err) {
    if (!rows_window_ok(window)) return FUSED_NA;
    const WinIdx w = fused_window(nsweep, tm, false, sl, window);
    if (w.out < tm.ratio || w.out < window / 2) return FUSED_NA;
    const int64_t grid = (n + w.out - 1) / w.out;
    if (partial && grid > partial_cap) return FUSED_NA;
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0):                                                  \
        if (nblocks) *nblocks = (int)grid;                                                               \
        if (window == 32)                                                                                \
            *err = launch_fused(r_up<MM, MCC, 32, SS, DG>, (unsigned)grid, MM * 32,                      \
                                sizeof(RowsSmem<MM, 32, SS, DG>), st, pdl, mat, d.ilo, d.iup, b, xin, xout, P0, \
                                P1, tm, xcoarse, n, alpha, nsweep, w, partial, sl);                      \
        else                                                                                             \
            *err = launch_fused(r_up<MM, MCC, 64, SS, DG>, (unsigned)grid, MM * 64,                      \
                                sizeof(RowsSmem<MM, 64, SS, DG>), st, pdl, mat, d.ilo, d.iup, b, xin, xout, P0, \
                                P1, tm, xcoarse, n, alpha, nsweep, w, partial, sl);                      \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;
        ROWS_COMBOS(X)
#undef X
        default: return FUSED_NA;
    }
}
