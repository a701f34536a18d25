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

// Block rows per thread (R).  With R = 1 the M row-threads of an element each read its M iterate values and
// (block smoothers) its M residual values and their M Dinv entries from shared memory in every sweep: ncu
// showed the L1 / shared-memory pipe at 82-86 % and the kernel at 0.57 of the HBM peak
// (profiles/r01c_ncu_summary.md).  With R = 2 / 3 a thread owns R consecutive rows: every iterate / residual
// value it fetches from shared memory feeds R rows, and the R rows of Dinv live in registers next to the R
// rows of A_di, so shared-memory reads per row and sweep drop from ~3 M + 1 to ~2 M / R.  An element then
// takes G = ceil(M / R) threads (rows >= M of the last thread are padding: zero operator, no stores).
//
// Launch bounds: R = 1 aims at 1152 resident threads per SM (<= 56 registers per thread) for the compressed
// structure classes and 768 (<= 80 registers) for dense off-diagonal blocks, whose rows of A_lo and A_up are
// register resident too (3 M doubles).  R > 1: registers per thread are estimated from the resident doubles.
#ifndef ROWS_THREADS_PER_SM
#define ROWS_THREADS_PER_SM 1152
#endif
constexpr int rows_group(int m, int r) { return (m + r - 1) / r; }
constexpr int rows_min_blocks(int m, int w, int st, bool diag, int r) {
    if (r == 1) {
        const int target = st == AMG1D_ST_DENSE ? 768 : ROWS_THREADS_PER_SM;
        return target / (m * w) >= 1 ? target / (m * w) : 1;
    }
    const int doubles = r * (m + (diag ? 1 : m) + (st == AMG1D_ST_DENSE ? 2 * m : 1));
    const int regs = 2 * doubles + 50;
    const int nb = 65536 / (regs * rows_group(m, r) * w);
    return nb < 1 ? 1 : (nb > 5 ? 5 : nb);
}

template <int M, int W, int ST, bool DIAG, int R>
struct RowsSmem {
    static constexpr int XS = W + 3;          // slots 0 and W + 1 are the (zero) window edges
    static constexpr int RS = W + 9;
    static constexpr int G = rows_group(M, R);            // threads per element
    static constexpr int NT = G * W;                      // threads per CTA
    static constexpr bool DVS = !DIAG && R == 1;          // Dinv rows in shared memory (R = 1), else registers
    double xs[2][M][XS];                      // iterate exchange, xs[buf][row][window position + 1]
    double rs[M][RS];                         // b stage, residual exchange, restriction source
    double dv[DVS ? M : 1][DVS ? M * W : 1];  // R = 1: dv[j][t] = Dinv[i, j] of thread t's element (block smoother)
    double sv[ST == AMG1D_ST_DENSE ? 1 : M][W];  // the one stored row of A_up (ST_COLROW) / A_lo (ST_ROWCOL)
};

// what a thread (rows g R .. g R + R - 1) keeps in registers
template <int M, int ST, bool DIAG, int R>
struct RowOp {
    static constexpr bool DVR = !DIAG && R > 1;
    double di[R][M];                                  // A_di[i, :]
    double lo[R][ST == AMG1D_ST_DENSE ? M : 1];       // A_lo[i, :]  |  ST_COLROW: A_lo[i, ilo]  |  ST_ROWCOL: unused
    double up[R][ST == AMG1D_ST_DENSE ? M : 1];       // A_up[i, :]  |  ST_ROWCOL: A_up[i, iup]  |  ST_COLROW: unused
    double dinv[R];                                   // point Jacobi: Dinv[i]
    double dvr[R][DVR ? M : 1];                       // block smoother, R > 1: Dinv[i, :]
};

// T points at the element's first stored entry, STRIDE = distance between consecutive tile rows k:
// AMG1D_TILE in the element tiles, 1 in a pattern table (PatOp)
template <int M, int W, int ST, bool DIAG, int R, int STRIDE>
__device__ __forceinline__ void rows_load_from(const double* __restrict__ T, bool active, int ilo,
                                               int iup, int g, int w, int t, RowOp<M, ST, DIAG, R>& A,
                                               RowsSmem<M, W, ST, DIAG, R>& S) {
    using O = OpShape<M, ST>;
    using SM = RowsSmem<M, W, ST, DIAG, R>;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = g * R + r;
        const bool on = active && (R == 1 || i < M);
        A.dinv[r] = 0.0;
        if (on) {
            if constexpr (DIAG) {
                A.dinv[r] = T[(O::O_DV + i) * STRIDE];
            } else if constexpr (SM::DVS) {
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    if constexpr (STRIDE == 1) S.dv[j][t] = T[O::O_DV + j * M + i];   // pattern table: broadcast loads
                    else cp_async8(&S.dv[j][t], T + (O::O_DV + j * M + i) * STRIDE);
                }
            } else {
#pragma unroll
                for (int j = 0; j < M; ++j) A.dvr[r][j] = T[(O::O_DV + j * M + i) * STRIDE];
            }
#pragma unroll
            for (int j = 0; j < M; ++j) A.di[r][j] = T[(O::O_DI + j * M + i) * STRIDE];
            if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
                for (int j = 0; j < M; ++j) A.lo[r][j] = T[(j * M + i) * STRIDE];
#pragma unroll
                for (int j = 0; j < M; ++j) A.up[r][j] = T[(O::O_UP + j * M + i) * STRIDE];
            } else if constexpr (ST == AMG1D_ST_COLROW) {
                A.lo[r][0] = T[i * STRIDE];
                A.up[r][0] = 0.0;
                if (i == iup) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        if constexpr (STRIDE == 1) S.sv[j][w] = T[O::O_UP + j];
                        else cp_async8(&S.sv[j][w], T + (O::O_UP + j) * STRIDE);
                    }
                }
            } else {
                A.up[r][0] = T[(O::O_UP + i) * STRIDE];
                A.lo[r][0] = 0.0;
                if (i == ilo) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        if constexpr (STRIDE == 1) S.sv[j][w] = T[j];
                        else cp_async8(&S.sv[j][w], T + j * STRIDE);
                    }
                }
            }
        } else {
            if constexpr (SM::DVS) {
#pragma unroll
                for (int j = 0; j < M; ++j) S.dv[j][t] = 0.0;
            } else if constexpr (!DIAG) {
#pragma unroll
                for (int j = 0; j < M; ++j) A.dvr[r][j] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) A.di[r][j] = 0.0;
            if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
                for (int j = 0; j < M; ++j) { A.lo[r][j] = 0.0; A.up[r][j] = 0.0; }
            } else {
                A.lo[r][0] = 0.0;
                A.up[r][0] = 0.0;
                if (i == (ST == AMG1D_ST_COLROW ? iup : ilo)) {     // inactive element: its stored row is zero
#pragma unroll
                    for (int j = 0; j < M; ++j) S.sv[j][w] = 0.0;
                }
            }
        }
    }
}

// e = local element index (element tiles), eg = global element index (pattern table of a translation-
// invariant level, PatOp in kernels_fused.cuh)
template <int M, int W, int ST, bool DIAG, int R>
__device__ __forceinline__ void rows_load(const double* __restrict__ mat, const PatOp& po, int64_t e, int64_t eg,
                                          bool active, int ilo, int iup, int g, int w, int t,
                                          RowOp<M, ST, DIAG, R>& A, RowsSmem<M, W, ST, DIAG, R>& S) {
    using O = OpShape<M, ST>;
    constexpr int K = O::O_DV + (DIAG ? M : M * M);
    if (po.tab != nullptr) {
        active = active && eg >= 0 && eg < po.n_glob;
        const int64_t s = !active ? 0 : eg < po.n_head ? eg
                        : (eg >= po.n_glob - po.n_tail ? po.n_head + 1 + (eg - (po.n_glob - po.n_tail)) : po.n_head);
        rows_load_from<M, W, ST, DIAG, R, 1>(po.tab + s * K, active, ilo, iup, g, w, t, A, S);
    } else {
        rows_load_from<M, W, ST, DIAG, R, AMG1D_TILE>(
            mat + (active ? (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31) : 0), active, ilo, iup, g, w, t, A, S);
    }
}

// rows g R .. g R + R - 1 of  A_lo x[w - 1] + A_di x[w] + A_up x[w + 1]  from exchange buffer `buf`; every row
// in the order of g_row_Ax (each shared-memory value is fetched once and feeds the R rows)
template <int M, int W, int ST, bool DIAG, int R>
__device__ __forceinline__ void rows_Ax(const RowOp<M, ST, DIAG, R>& A, const RowsSmem<M, W, ST, DIAG, R>& S,
                                        int buf, int ilo, int iup, int g, int w, double (&y)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) y[r] = 0.0;
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const double xv = S.xs[buf][j][w];
#pragma unroll
            for (int r = 0; r < R; ++r) y[r] = fma(A.lo[r][j], xv, y[r]);
        }
    } else if constexpr (ST == AMG1D_ST_COLROW) {
        const double xv = S.xs[buf][ilo][w];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = fma(A.lo[r][0], xv, 0.0);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (g * R + r == ilo) {
#pragma unroll
                for (int j = 0; j < M; ++j) y[r] = fma(S.sv[j][w], S.xs[buf][j][w], y[r]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double xv = S.xs[buf][j][w + 1];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = fma(A.di[r][j], xv, y[r]);
    }
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const double xv = S.xs[buf][j][w + 2];
#pragma unroll
            for (int r = 0; r < R; ++r) y[r] = fma(A.up[r][j], xv, y[r]);
        }
    } else if constexpr (ST == AMG1D_ST_COLROW) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (g * R + r == iup) {
#pragma unroll
                for (int j = 0; j < M; ++j) y[r] = fma(S.sv[j][w], S.xs[buf][j][w + 2], y[r]);
            }
        }
    } else {
        const double xv = S.xs[buf][iup][w + 2];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = fma(A.up[r][0], xv, y[r]);
    }
}

// x_i += alpha (Dinv r)_i; block smoothers exchange r among the threads of the element (one barrier)
template <int M, int W, int ST, bool DIAG, int R>
__device__ __forceinline__ void rows_update(const RowOp<M, ST, DIAG, R>& A, RowsSmem<M, W, ST, DIAG, R>& S,
                                            const double (&res)[R], double (&xc)[R], double alpha, int g, int w,
                                            int t) {
    using SM = RowsSmem<M, W, ST, DIAG, R>;
    if constexpr (DIAG) {
#pragma unroll
        for (int r = 0; r < R; ++r) xc[r] = __dadd_rn(xc[r], __dmul_rn(alpha, A.dinv[r] * res[r]));
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (R == 1 || g * R + r < M) S.rs[g * R + r][w] = res[r];
        __syncthreads();
        double z[R];
#pragma unroll
        for (int r = 0; r < R; ++r) z[r] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const double rv = S.rs[j][w];
            if constexpr (SM::DVS) {
                z[0] = fma(S.dv[j][t], rv, z[0]);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) z[r] = fma(A.dvr[r][j], rv, z[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) xc[r] = __dadd_rn(xc[r], __dmul_rn(alpha, z[r]));
    }
}

template <int M, int W, int ST, bool DIAG, int R>
__device__ __forceinline__ void rows_init_edges(RowsSmem<M, W, ST, DIAG, R>& S) {
    static_assert(RowsSmem<M, W, ST, DIAG, R>::NT >= 4 * M, "CTA too small to clear the window edges");
    if (threadIdx.x < 4 * M) {
        const int buf = threadIdx.x / (2 * M);
        const int i = (threadIdx.x >> 1) % M;
        S.xs[buf][i][(threadIdx.x & 1) ? W + 1 : 0] = 0.0;
    }
}

#define ROWS_BOUNDS __launch_bounds__(rows_group(M, R) * W, rows_min_blocks(M, W, ST, DIAG, R))

// nsweep pre-smoothing sweeps + residual + restriction (f_down for large blocks)
template <int M, int MC, int W, int ST, bool DIAG, int R>
__global__ void ROWS_BOUNDS
r_down(const double* __restrict__ mat, PatOp po, int ilo, int iup, const double* __restrict__ b,
       const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
       const double* __restrict__ P1, TransferMap tm, double* __restrict__ rc, int64_t n, double alpha,
       int nsweep, int zero_guess, WinIdx wi, Slab sl) {
    using SM = RowsSmem<M, W, ST, DIAG, R>;
    extern __shared__ __align__(16) double rows_smem[];
    SM& S = *reinterpret_cast<SM*>(rows_smem);
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int w = t % W, g = t / W;
    const int halo = wi.halo, out = wi.out;
    const int64_t e0 = (int64_t)blockIdx.x * out - halo;          // local index of window element 0
    const int64_t e = e0 + w;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    rows_init_edges<M, W, ST, DIAG, R>(S);
    RowOp<M, ST, DIAG, R> A;
    rows_load<M, W, ST, DIAG, R>(mat, po, e, e + sl.e_off, active, ilo, iup, g, w, t, A, S);   // independent of earlier kernels
    pdl_wait();
    // coalesced mapping: thread t moves doubles t, t + NT, ... of the window's contiguous range
#pragma unroll
    for (int q = t; q < M * W; q += SM::NT) {
        const int we = q / M, ie = q - we * M;
        const int64_t ee = e0 + we;
        const bool act = ee >= -(int64_t)sl.gl && ee < n + sl.gr;
        S.rs[ie][we] = act ? b[e0 * M + q] : 0.0;
        S.xs[0][ie][we + 1] = (act && !zero_guess) ? xin[e0 * M + q] : 0.0;
    }
    cp_async_commit_wait_all();
    __syncthreads();
    double bb[R], xc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool row = R == 1 || g * R + r < M;
        bb[r] = row ? S.rs[g * R + r][w] : 0.0;
        xc[r] = row ? S.xs[0][g * R + r][w + 1] : 0.0;
    }
    int buf = 0;
    double res[R];
    for (int s = 0; s < nsweep; ++s) {
        if (zero_guess && s == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) res[r] = bb[r] - 0.0;
        } else {
            rows_Ax<M, W, ST, DIAG, R>(A, S, buf, ilo, iup, g, w, res);
#pragma unroll
            for (int r = 0; r < R; ++r) res[r] = bb[r] - res[r];
        }
        rows_update<M, W, ST, DIAG, R>(A, S, res, xc, alpha, g, w, t);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (R == 1 || g * R + r < M) S.xs[buf ^ 1][g * R + r][w + 1] = xc[r];
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int q = t; q < M * W; q += SM::NT) {
        const int we = q / M, ie = q - we * M;
        if (we >= halo && we < halo + out && e0 + we < n) xout[e0 * M + q] = S.xs[buf][ie][we + 1];
    }
    // residual with the final iterate, then restriction (order of g_restrict)
    rows_Ax<M, W, ST, DIAG, R>(A, S, buf, ilo, iup, g, w, res);
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (R == 1 || g * R + r < M) S.rs[g * R + r][w] = bb[r] - res[r];
    __syncthreads();
    const bool mine = w >= halo && w < halo + out;
    const int64_t eg = e + sl.e_off;
    if (g * R < MC && mine && eg < tm.n_fine) {
        int kdiv, kmod;
        small_divmod(wi.qmod0 + w, tm.ratio, &kdiv, &kmod);
        if (kmod == 0) {
            const int64_t Kc = (int64_t)blockIdx.x * wi.opr + wi.qdiv0 + kdiv + tm.base;   // eg == tm.first(Kc)
            const int64_t Kl = Kc - sl.c_off;
            if (Kc >= 0 && Kc < tm.n_coarse && Kl >= 0 && Kl < sl.nc) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = g * R + r;
                    if (i >= MC) break;
                    double acc = 0.0;
                    if (P1) {                                   // children of Kc - 1: [eg - ratio, eg), clamped
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
}

// prolongation + correction + nsweep post-smoothing sweeps (+ || b - A x ||^2 partial sums)
template <int M, int MC, int W, int ST, bool DIAG, int R>
__global__ void ROWS_BOUNDS
r_up(const double* __restrict__ mat, PatOp po, int ilo, int iup, const double* __restrict__ b,
     const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
     const double* __restrict__ P1, TransferMap tm, const double* __restrict__ xcoarse, int64_t n,
     double alpha, int nsweep, WinIdx wi, double* __restrict__ partial, Slab sl) {
    using SM = RowsSmem<M, W, ST, DIAG, R>;
    extern __shared__ __align__(16) double rows_smem[];
    SM& S = *reinterpret_cast<SM*>(rows_smem);
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int w = t % W, g = t / W;
    const int halo = wi.halo, out = wi.out;
    const int64_t e0 = (int64_t)blockIdx.x * out - halo;
    const int64_t e = e0 + w;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    rows_init_edges<M, W, ST, DIAG, R>(S);
    RowOp<M, ST, DIAG, R> A;
    rows_load<M, W, ST, DIAG, R>(mat, po, e, e + sl.e_off, active, ilo, iup, g, w, t, A, S);
    pdl_wait();
#pragma unroll
    for (int q = t; q < M * W; q += SM::NT) {
        const int we = q / M, ie = q - we * M;
        const int64_t ee = e0 + we;
        double bv = 0.0, xv = 0.0;
        if (ee >= -(int64_t)sl.gl && ee < n + sl.gr) {
            bv = b[e0 * M + q];
            xv = xin[e0 * M + q];
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
    double bb[R], xc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool row = R == 1 || g * R + r < M;
        bb[r] = row ? S.rs[g * R + r][w] : 0.0;
        xc[r] = row ? S.xs[0][g * R + r][w + 1] : 0.0;
    }
    int buf = 0;
    double res[R];
    for (int s = 0; s < nsweep; ++s) {
        rows_Ax<M, W, ST, DIAG, R>(A, S, buf, ilo, iup, g, w, res);
#pragma unroll
        for (int r = 0; r < R; ++r) res[r] = bb[r] - res[r];
        rows_update<M, W, ST, DIAG, R>(A, S, res, xc, alpha, g, w, t);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (R == 1 || g * R + r < M) S.xs[buf ^ 1][g * R + r][w + 1] = xc[r];
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int q = t; q < M * W; q += SM::NT) {
        const int we = q / M, ie = q - we * M;
        if (we >= halo && we < halo + out && e0 + we < n) xout[e0 * M + q] = S.xs[buf][ie][we + 1];
    }
    if (partial) {
        rows_Ax<M, W, ST, DIAG, R>(A, S, buf, ilo, iup, g, w, res);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (R == 1 || g * R + r < M) S.rs[g * R + r][w] = bb[r] - res[r];
        __syncthreads();
        double s2 = 0.0;
        if (g == 0 && e < n && w >= halo && w < halo + out) {
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

// (window, rows per thread) variants that are compiled
#define ROWS_VARIANTS(V) V(32, 1) V(64, 1) V(32, 2) V(64, 2) V(32, 3) V(64, 3)

inline bool rows_window_ok(int window) { return window == 32 || window == 64; }
inline bool rows_rpt_ok(int rpt) { return rpt >= 1 && rpt <= 3; }

template <int M, int MC, int W, int ST, bool DIAG, int R>
inline cudaError_t rows_configure_t() {
    const int smem = (int)sizeof(RowsSmem<M, W, ST, DIAG, R>);
    cudaError_t e = cudaFuncSetAttribute(r_down<M, MC, W, ST, DIAG, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(r_up<M, MC, W, ST, DIAG, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// once per device context and level shape, before the first launch (amg1d_finalize); *have = a kernel exists
inline cudaError_t rows_configure(const MatDesc& d, int mc, bool* have) {
    *have = true;
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define V(WW, RR)                                                                                        \
    if (e == cudaSuccess) e = rows_configure_t<RM, RMC, WW, RST, RDG, RR>();
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0): {                                                \
        constexpr int RM = MM, RMC = MCC, RST = SS;                                                      \
        constexpr bool RDG = DG;                                                                         \
        cudaError_t e = cudaSuccess;                                                                     \
        ROWS_VARIANTS(V)                                                                                 \
        return e;                                                                                        \
    }
        ROWS_COMBOS(X)
#undef X
#undef V
        default: *have = false; return cudaSuccess;
    }
}

// Same contract as fused_down / fused_up (FUSED_NA / FUSED_OK / FUSED_ERR).
inline int rows_down(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, bool zero,
                     const double* mat, const PatOp& po, const double* b, const double* xin, double* xout,
                     const double* P0, const double* P1, double* rc, int64_t n, int64_t n_cover,
                     double alpha, const Slab& sl, int window, int rpt, cudaStream_t st, bool pdl,
                     cudaError_t* err) {
    if (!rows_window_ok(window) || !rows_rpt_ok(rpt)) return FUSED_NA;
    const WinIdx w = fused_window(nsweep, tm, P1 != nullptr || tm.shift != 0 || tm.base != 0, sl, window);
    if (w.out < tm.ratio || w.out < window / 2) return FUSED_NA;
    const unsigned grid = (unsigned)((n_cover + w.out - 1) / w.out);
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define V(WW, RR)                                                                                        \
    if (window == WW && rpt == RR)                                                                       \
        *err = launch_fused(r_down<RM, RMC, WW, RST, RDG, RR>, grid, RowsSmem<RM, WW, RST, RDG, RR>::NT, \
                            sizeof(RowsSmem<RM, WW, RST, RDG, RR>), st, pdl, mat, po, d.ilo, d.iup, b,   \
                            xin, xout, P0, P1, tm, rc, n, alpha, nsweep, zero ? 1 : 0, w, sl);
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0): {                                                \
        constexpr int RM = MM, RMC = MCC, RST = SS;                                                      \
        constexpr bool RDG = DG;                                                                         \
        ROWS_VARIANTS(V)                                                                                 \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;                                               \
    }
        ROWS_COMBOS(X)
#undef X
#undef V
        default: return FUSED_NA;
    }
}

inline int rows_up(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, const double* mat,
                   const PatOp& po, const double* b, const double* xin, double* xout, const double* P0, const double* P1,
                   const double* xcoarse, int64_t n, double alpha, double* partial, int64_t partial_cap,
                   int* nblocks, const Slab& sl, int window, int rpt, cudaStream_t st, bool pdl,
                   cudaError_t* err) {
    if (!rows_window_ok(window) || !rows_rpt_ok(rpt)) return FUSED_NA;
    const WinIdx w = fused_window(nsweep, tm, false, sl, window);
    if (w.out < tm.ratio || w.out < window / 2) return FUSED_NA;
    const int64_t grid = (n + w.out - 1) / w.out;
    if (partial && grid > partial_cap) return FUSED_NA;
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define V(WW, RR)                                                                                        \
    if (window == WW && rpt == RR)                                                                       \
        *err = launch_fused(r_up<RM, RMC, WW, RST, RDG, RR>, (unsigned)grid,                             \
                            RowsSmem<RM, WW, RST, RDG, RR>::NT, sizeof(RowsSmem<RM, WW, RST, RDG, RR>),  \
                            st, pdl, mat, po, d.ilo, d.iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha, \
                            nsweep, w, partial, sl);
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0): {                                                \
        constexpr int RM = MM, RMC = MCC, RST = SS;                                                      \
        constexpr bool RDG = DG;                                                                         \
        if (nblocks) *nblocks = (int)grid;                                                               \
        ROWS_VARIANTS(V)                                                                                 \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;                                               \
    }
        ROWS_COMBOS(X)
#undef X
#undef V
        default: return FUSED_NA;
    }
}
