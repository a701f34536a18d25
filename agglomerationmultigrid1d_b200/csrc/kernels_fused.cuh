// Fast kernel tier: templated on the block size M and the level's structure class ST (layout.cuh),
// one thread per element, operator read straight from the element-tile layout with fully coalesced
// 256-byte warp requests.
//
//  f_sweep / f_resnorm / f_residual_restrict  - streaming kernels (one pass over the operator each)
//  f_down   nPre sweeps + residual + restriction  in ONE pass over the level's operator
//  f_up     prolongation + correction + nPost sweeps (+ optional ||b - A x||^2)  in ONE pass
//  f_tail   the whole sub-V-cycle of the coarse levels (<= 512 elements) in ONE CTA, out of shared memory
//
// f_down / f_up keep the element's operator blocks in registers for the whole leg (Dinv in shared
// memory, staged with cp.async) and exchange only the M iterate values with the two neighbour
// threads through shared memory between sweeps (Jacobi needs the *old* neighbour values, so the
// exchange is double buffered).  A CTA owns a window of B consecutive elements; after s sweeps only
// the inner [s, B - s) elements are still exact, so the CTA emits the inner B - 2 halo elements and
// adjacent CTAs overlap by the halo (the overlap is re-read through L2, not HBM).  Per element the
// arithmetic and its order are identical to the generic tier, so both tiers agree bit for bit on x.
//
// Both smoother kinds (block Jacobi: Dinv = m x m; point Jacobi: Dinv = m) and both transfer kinds
// (single parent: dg_dg / aggdg_dg / aggdg_aggdg; two parents: cg_cg / dg_cg / aggdg_cg) are covered,
// the latter with the closed-form parent map  parent(e) = (e + shift) / ratio + base.
//
// Algorithmic bytes per element (FP64; Kop = doubles of the level's structure class, layout.cuh;
// coarse block size MC, ratio R children):
//   f_sweep               8 (Kop + 3 M)                           [zero guess: 8 (|Dinv| + 2 M)]
//   f_down (S sweeps)     8 (Kop + 3 M + MC / R)  (+ P block if it is per-element)
//   f_up   (S sweeps)     8 (Kop + 3 M + MC / R)  (+ P block if it is per-element)
// against S * 8 (4 M^2 + 3 M) + 8 (3 M^2 + 2 M + ...) for the unfused dense sequence (SURVEY 8d B_ref).
#pragma once
#include <cuda_runtime.h>
#include "kernels_generic.cuh"
#include "layout.cuh"
#include "halo_p2p.cuh"

// Position of a rank's slab inside its level (single GPU: gl = gr = e_off = c_off = 0).  Local element
// index e runs over [-gl, n + gr): gl / gr ghost elements (with operator blocks, rhs and iterate) on the
// left / right slab edge; e_off = global index of local element 0; c_off = global index of local coarse
// element 0; nc = coarse elements this rank owns.
struct Slab {
    int gl, gr;
    int64_t e_off, c_off, nc;
};

// Translation-invariant level (uniform mesh, amg1d_set_level_pattern) in "pattern-resident" form: the
// n_head + 1 + n_tail distinct block sets of the level as a small table tab[set][k] (k = tile row of
// layout.cuh), instead of one stored block set per element.  The fused legs then fetch an element's
// operator from the table (a few KB, served by L1 as a warp-wide broadcast) and HBM carries only the
// vectors.  tab == nullptr: the element tiles are streamed (the general per-element layout).
struct PatOp {
    const double* tab;
    int64_t n_glob;        // elements of the whole level (the table is indexed by GLOBAL element)
    int n_head, n_tail;
    const double* host_interior;   // HOST copy of the interior row tab[n_head][*] (pattern_resident = 2: the
                                   // launcher passes it by value as ParamOp), else nullptr; unused on the device
};

// Per-launch constants of a fused leg, computed on the host so that the per-thread index arithmetic
// (parent element, owner test, transfer-block index) is 32-bit: with out % ratio == 0 the CTA's first
// thread has (e_global + shift) = ratio * (blockIdx * opr + qdiv0) + qmod0.
struct WinIdx {
    int halo, out;     // recomputed window elements per side; elements emitted per CTA
    int opr;           // out / ratio
    int qmod0;         // floormod(e_off + shift - halo, ratio)
    int64_t qdiv0;     // floordiv(e_off + shift - halo, ratio)
    int pmod0;         // floormod(e_off - halo - n_head, period) + pbias, or -1: use TransferMap::blk (64-bit)
};

__device__ __forceinline__ void small_divmod(int x, int d, int* q, int* r) {   // x >= 0, d >= 1
    if (d == 1) { *q = x; *r = 0; }
    else if (d == 2) { *q = x >> 1; *r = x & 1; }
    else { *q = x / d; *r = x - *q * d; }
}

// transfer-block index of global fine element c = (window thread tc of this CTA), see TransferMap::blk
__device__ __forceinline__ int64_t win_blk(const TransferMap& tm, const WinIdx& w, int64_t c, int tc) {
    if (tm.period == 0 || w.pmod0 < 0) return tm.blk(c);
    if (c < tm.n_head) return c;
    if (c >= tm.n_fine - tm.n_tail) return tm.n_head + tm.period + (c - (tm.n_fine - tm.n_tail));
    int q, r;
    small_divmod(w.pmod0 + tc, tm.period, &q, &r);
    return tm.n_head + r;
}

// Programmatic dependent launch (PDL): every fused kernel first lets its successor in the stream start
// launching, issues the loads that do not depend on earlier kernels (the level's operator - most of its
// bytes), and only then waits for the preceding kernels to complete.  The successor's CTAs fill the SMs
// that the last wave of this grid leaves idle, and the launch latency of the small levels disappears.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_fused(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifndef FUSED_B
#define FUSED_B 128  // window (threads) per CTA of f_down / f_up
#endif
#ifndef FUSED_MINB
#define FUSED_MINB 3   // __launch_bounds__ min CTAs per SM for f_down / f_up
#endif
#define TAIL_B 512     // threads (= max elements of the first tail level) of f_tail

// ---- small helpers ---------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void load_vec(const double* __restrict__ p, double (&v)[M]) {
    if constexpr (M % 2 == 0) {
#pragma unroll
        for (int i = 0; i < M; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + i);
            v[i] = t.x;
            v[i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) v[i] = p[i];
    }
}

template <int M>
__device__ __forceinline__ void store_vec(double* __restrict__ p, const double (&v)[M]) {
    if constexpr (M % 2 == 0) {
#pragma unroll
        for (int i = 0; i < M; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) p[i] = v[i];
    }
}

// The fast tier handles ST_COLROW levels whose A_up row is the reference's DG trace row (local node 1,
// src/dg_mesh.jl:41-46: left end, right end, interior nodes), so that the row is a compile-time
// register; any other row index stays with the generic tier (fast_tier_ok below).
#define FUSED_COLROW_IUP 1

// Tile rows of the level's operator part per structure class.
template <int M, int ST>
struct OpShape {
    static constexpr int NO = ST ? M : M * M;           // stored entries of A_lo (and of A_up)
    static constexpr int O_DI = NO, O_UP = NO + M * M, O_DV = 2 * NO + M * M;
};

// y = A_lo xl + A_di xc + A_up xr with the operator streamed from the tile (T points at
// [tile][0][lane]).  Neighbour values: ST_COLROW needs only xl[ilo], passed as xl[0]; ST_ROWCOL needs
// only xr[iup], passed as xr[0].  Same accumulation order as g_row_Ax.
template <int M, int ST>
__device__ __forceinline__ void stream_Ax(const double* __restrict__ T, int ilo, int iup,
                                          const double (&xl)[M], const double (&xc)[M],
                                          const double (&xr)[M], double (&y)[M]) {
    using S = OpShape<M, ST>;
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(T[(j * M + i) * AMG1D_TILE], xl[j], y[i]);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[i * AMG1D_TILE], xl[0], 0.0);
    } else {
        double yr = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) yr = fma(T[j * AMG1D_TILE], xl[j], yr);
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = (i == ilo) ? yr : 0.0;
    }
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[(S::O_DI + j * M + i) * AMG1D_TILE], xc[j], y[i]);
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(T[(S::O_UP + j * M + i) * AMG1D_TILE], xr[j], y[i]);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
        double yr = y[FUSED_COLROW_IUP];
#pragma unroll
        for (int j = 0; j < M; ++j) yr = fma(T[(S::O_UP + j) * AMG1D_TILE], xr[j], yr);
        y[FUSED_COLROW_IUP] = yr;
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[(S::O_UP + i) * AMG1D_TILE], xr[0], y[i]);
    }
}

// Neighbour blocks of element e from a global vector, in the form stream_Ax / reg_Ax expect.
template <int M, int ST>
__device__ __forceinline__ void load_neighbours(const double* __restrict__ x, int64_t e, int ilo, int iup,
                                                double (&xl)[M], double (&xr)[M]) {
    if constexpr (ST == AMG1D_ST_COLROW) xl[0] = x[(e - 1) * M + ilo];
    else load_vec<M>(x + (e - 1) * M, xl);
    if constexpr (ST == AMG1D_ST_ROWCOL) xr[0] = x[(e + 1) * M + iup];
    else load_vec<M>(x + (e + 1) * M, xr);
}

// ---- streaming kernels --------------------------------------------------------------------------------
template <int M, bool DIAG, int ST>
__global__ void __launch_bounds__(256) f_sweep(const double* __restrict__ mat, int ilo, int iup,
                                               const double* __restrict__ b,
                                               const double* __restrict__ xin,
                                               double* __restrict__ xout, int64_t n, double alpha,
                                               int zero_guess) {
    using S = OpShape<M, ST>;
    constexpr int K = S::O_DV + (DIAG ? M : M * M);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double* T = mat + (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31);
    double r[M], xc[M], bb[M];
    load_vec<M>(b + e * M, bb);
    if (zero_guess) {
#pragma unroll
        for (int i = 0; i < M; ++i) { xc[i] = 0.0; r[i] = bb[i] - 0.0; }
    } else {
        double xl[M], xr[M], y[M];
        load_vec<M>(xin + e * M, xc);
        load_neighbours<M, ST>(xin, e, ilo, iup, xl, xr);
        stream_Ax<M, ST>(T, ilo, iup, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
    }
    double xn[M];
    if constexpr (DIAG) {
#pragma unroll
        for (int i = 0; i < M; ++i)
            xn[i] = __dadd_rn(xc[i], __dmul_rn(alpha, T[(S::O_DV + i) * AMG1D_TILE] * r[i]));
    } else {
        double z[M];
#pragma unroll
        for (int i = 0; i < M; ++i) z[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) z[i] = fma(T[(S::O_DV + j * M + i) * AMG1D_TILE], r[j], z[i]);
#pragma unroll
        for (int i = 0; i < M; ++i) xn[i] = __dadd_rn(xc[i], __dmul_rn(alpha, z[i]));
    }
    store_vec<M>(xout + e * M, xn);
}

// partial[blockIdx] = sum over the block's elements of || b - A x ||^2
template <int M, int ST>
__global__ void __launch_bounds__(256) f_resnorm(const double* __restrict__ mat, int K, int ilo, int iup,
                                                 const double* __restrict__ b,
                                                 const double* __restrict__ x, int64_t n,
                                                 double* __restrict__ partial) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (e < n) {
        const double* T = mat + (e >> 5) * (int64_t)K * AMG1D_TILE + (e & 31);
        double xl[M], xc[M], xr[M], y[M], bb[M];
        load_vec<M>(b + e * M, bb);
        load_vec<M>(x + e * M, xc);
        load_neighbours<M, ST>(x, e, ilo, iup, xl, xr);
        stream_Ax<M, ST>(T, ilo, iup, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) { const double r = bb[i] - y[i]; s = fma(r, r, s); }
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// y = A x and partial[blockIdx] = sum over the block's elements of x . y   (the p' A p of conjugate
// gradients fused into the product; partial may be null)
template <int M, int ST>
__global__ void __launch_bounds__(256) f_matvec_dot(const double* __restrict__ mat, int K, int ilo, int iup,
                                                    const double* __restrict__ x, double* __restrict__ y,
                                                    int64_t n, double* __restrict__ partial) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (e < n) {
        const double* T = mat + (e >> 5) * (int64_t)K * AMG1D_TILE + (e & 31);
        double xl[M], xc[M], xr[M], yy[M];
        load_vec<M>(x + e * M, xc);
        load_neighbours<M, ST>(x, e, ilo, iup, xl, xr);
        stream_Ax<M, ST>(T, ilo, iup, xl, xc, xr, yy);
        store_vec<M>(y + e * M, yy);
#pragma unroll
        for (int i = 0; i < M; ++i) s = fma(xc[i], yy[i], s);
    }
    if (partial) {
        s = block_sum(s);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

// ---- register-resident multi-sweep kernels -----------------------------------------------------------
// Shared-memory exchange buffers, structure-of-arrays so that neighbouring threads hit neighbouring
// banks: xs[buf][i][slot], slot = thread + 1, slots 0 and B+1 stay zero.
template <int M, int B>
struct Exchange {
    double xs[2][M][B + 2];
};

template <int M, int B>
__device__ __forceinline__ void exch_init(Exchange<M, B>& ex) {
    if (threadIdx.x < 2 * M * 2) {
        const int buf = threadIdx.x / (2 * M);
        const int i = (threadIdx.x / 2) % M;
        const int side = threadIdx.x & 1;
        ex.xs[buf][i][side ? B + 1 : 0] = 0.0;
    }
}

// The element's A_lo / A_di / A_up in registers (only the stored entries of the structure class).
template <int M, int ST>
struct RegOp {
    static constexpr bool has_dv = false;
    double lo[OpShape<M, ST>::NO], di[M * M], up[OpShape<M, ST>::NO];
};

// One block set (A_lo, A_di, A_up, Dinv in tile-row order = one row of a pattern table) passed BY VALUE as a
// __grid_constant__ kernel parameter: the entries are constant-bank operands of the FMAs - no loads, no
// registers, no shared memory for the operator.  Used by f_down_c / f_up_c for the CTAs whose whole window
// lies in the translation-invariant interior of a level (option pattern_resident = 2).
template <int M, int ST, bool DIAG>
struct ParamOp {
    static constexpr bool has_dv = true;
    double lo[OpShape<M, ST>::NO], di[M * M], up[OpShape<M, ST>::NO], dv[DIAG ? M : M * M];
};

template <int M, int ST, class OP>
__device__ __forceinline__ void reg_Ax(const OP& A, int ilo, int iup, const double (&xl)[M],
                                       const double (&xc)[M], const double (&xr)[M], double (&y)[M]) {
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(A.lo[j * M + i], xl[j], y[i]);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A.lo[i], xl[0], 0.0);
    } else {
        double yr = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) yr = fma(A.lo[j], xl[j], yr);
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = (i == ilo) ? yr : 0.0;
    }
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A.di[j * M + i], xc[j], y[i]);
    if constexpr (ST == AMG1D_ST_DENSE) {
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(A.up[j * M + i], xr[j], y[i]);
    } else if constexpr (ST == AMG1D_ST_COLROW) {
        double yr = y[FUSED_COLROW_IUP];
#pragma unroll
        for (int j = 0; j < M; ++j) yr = fma(A.up[j], xr[j], yr);
        y[FUSED_COLROW_IUP] = yr;
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A.up[i], xr[0], y[i]);
    }
}

// one damped (block-)Jacobi sweep on register-resident blocks; same operation order as g_sweep.
// Dinv is read from shared memory: dcol points at this thread's column of ds[k][thread], stride DS.
template <int M, int ST, bool DIAG, int DS, class OP>
__device__ __forceinline__ void reg_sweep(const OP& A, int ilo, int iup,
                                          const double* __restrict__ dcol, const double (&bb)[M],
                                          const double (&xl)[M], double (&xc)[M], const double (&xr)[M],
                                          double alpha, bool zero_guess) {
    double r[M];
    if (zero_guess) {
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - 0.0;
    } else {
        double y[M];
        reg_Ax<M, ST, OP>(A, ilo, iup, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
    }
    if constexpr (DIAG) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
            double dvi;
            if constexpr (OP::has_dv) dvi = A.dv[i]; else dvi = dcol[i * DS];
            xc[i] = __dadd_rn(xc[i], __dmul_rn(alpha, dvi * r[i]));
        }
    } else {
        double z[M];
#pragma unroll
        for (int i = 0; i < M; ++i) z[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double dvk;
                if constexpr (OP::has_dv) dvk = A.dv[j * M + i]; else dvk = dcol[(j * M + i) * DS];
                z[i] = fma(dvk, r[j], z[i]);
            }
#pragma unroll
        for (int i = 0; i < M; ++i) xc[i] = __dadd_rn(xc[i], __dmul_rn(alpha, z[i]));
    }
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// In-register inverse of a dense M x M block (column-major: entry (r, c) at a[c * M + r]): Gauss-Jordan with
// partial pivoting, statement for statement the arithmetic of k_dinv_recompute (device_setup.cuh) so that the
// result equals the stored inverse bit for bit.  The pivot search is a chain of compare-and-swap steps (row r
// is swapped up whenever |a(r, c)| exceeds the current pivot candidate, which leaves the largest entry on the
// diagonal) recorded as booleans: every array index is a compile-time constant and the swaps are selects, so
// the block never leaves the register file (an integer pivot index made the compiler index local memory).
// PIV = false: the level's set-up pass found that no element ever swaps (every pivot already sits on the diagonal,
// as for the symmetric positive definite, diagonally heavy blocks of the DG operators) - the compare-and-swap
// chain is skipped, which leaves exactly the same arithmetic at a third of the instructions.
template <int M, bool PIV>
__device__ __forceinline__ void reg_invert(double (&a)[M * M]) {
    bool sw[M][M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
        if constexpr (PIV) {
#pragma unroll
        for (int r = c + 1; r < M; ++r) {
            const bool s = fabs(a[c * M + r]) > fabs(a[c * M + c]);
            sw[c][r] = s;
#pragma unroll
            for (int q = 0; q < M; ++q) {
                const double x = a[q * M + c], y = a[q * M + r];
                a[q * M + c] = s ? y : x;
                a[q * M + r] = s ? x : y;
            }
        }
        }
        const double dd = 1.0 / a[c * M + c];
        a[c * M + c] = 1.0;
#pragma unroll
        for (int q = 0; q < M; ++q) a[q * M + c] *= dd;
#pragma unroll
        for (int r = 0; r < M; ++r) {
            if (r == c) continue;
            const double f = a[c * M + r];
            a[c * M + r] = 0.0;
#pragma unroll
            for (int q = 0; q < M; ++q) a[q * M + r] = fma(-f, a[q * M + c], a[q * M + r]);
        }
    }
    if constexpr (PIV) {
#pragma unroll
    for (int c = M - 1; c >= 0; --c) {            // undo the row swaps as column swaps, in reverse order
#pragma unroll
        for (int r2 = M - 1; r2 > c; --r2) {
            const bool s = sw[c][r2];
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const double x = a[c * M + r], y = a[r2 * M + r];
                a[c * M + r] = s ? y : x;
                a[r2 * M + r] = s ? x : y;
            }
        }
    }
    }
}

// Option recompute_dinv: this thread's Dinv = inv(A_di) from the registers into its shared-memory column
// (rec = 1: with the pivot chain, 2: the level never pivots).  Called right after the operator loads, BEFORE the
// programmatic-dependency wait and the vector loads: placing it after them (to overlap their latency) was
// measured 40 % slower on T level 0 - the vectors' registers are then live across the inversion and it spills.
template <int M, int B, int ST, bool DIAG>
__device__ __forceinline__ void recompute_dinv(const RegOp<M, ST>& A, double (*ds)[B], bool active, int rec) {
    if constexpr (!DIAG) {
        if (rec && active) {
            double w[M * M];
#pragma unroll
            for (int k = 0; k < M * M; ++k) w[k] = A.di[k];
            if (rec == 2) reg_invert<M, false>(w); else reg_invert<M, true>(w);
#pragma unroll
            for (int k = 0; k < M * M; ++k) ds[k][threadIdx.x] = w[k];
        }
    }
}
template <int M, int B, int ST, bool DIAG>
__device__ __forceinline__ void recompute_dinv(const ParamOp<M, ST, DIAG>&, double (*)[B], bool, int) {}

template <int M, int ST, class OP>
__device__ __forceinline__ void reg_residual(const OP& A, int ilo, int iup,
                                             const double (&bb)[M], const double (&xl)[M],
                                             const double (&xc)[M], const double (&xr)[M],
                                             double (&r)[M]) {
    double y[M];
    reg_Ax<M, ST, OP>(A, ilo, iup, xl, xc, xr, y);
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
}

// publish x into exchange buffer `buf`, barrier, fetch what the structure class needs of both neighbours
template <int M, int B, int ST>
__device__ __forceinline__ void exchange(Exchange<M, B>& ex, int buf, int ilo, int iup,
                                         const double (&xc)[M], double (&xl)[M], double (&xr)[M]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < M; ++i) ex.xs[buf][i][t + 1] = xc[i];
    __syncthreads();
    if constexpr (ST == AMG1D_ST_COLROW) {
        xl[0] = ex.xs[buf][ilo][t];
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) xl[i] = ex.xs[buf][i][t];
    }
    if constexpr (ST == AMG1D_ST_ROWCOL) {
        xr[0] = ex.xs[buf][iup][t + 2];
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) xr[i] = ex.xs[buf][i][t + 2];
    }
}

// A_lo / A_di / A_up go to registers; Dinv (used once per sweep) is copied global -> shared with
// cp.async, i.e. without register staging, into this thread's own column ds[k][thread].
// rec (block smoothers, streamed tiles only): Dinv is NOT loaded - the leg inverts A.di in registers
// (recompute_dinv) and puts the result into the same shared-memory column.
template <int M, int B, int ST, bool DIAG, int STRIDE>
__device__ __forceinline__ void load_blocks_from(const double* __restrict__ T, RegOp<M, ST>& A, double (*ds)[B],
                                                 bool rec = false) {
    using S = OpShape<M, ST>;
    constexpr int ND = DIAG ? M : M * M;
    const int t = threadIdx.x;
    if constexpr (STRIDE == 1) {   // pattern table: every lane reads the same address - plain loads broadcast
#pragma unroll                     // (cp.async with one source address per warp was measured far slower)
        for (int k = 0; k < ND; ++k) ds[k][t] = T[S::O_DV + k];
    } else {
        if (!rec) {
#pragma unroll
            for (int k = 0; k < ND; ++k) cp_async8(&ds[k][t], T + (S::O_DV + k) * STRIDE);
        }
    }
#pragma unroll
    for (int k = 0; k < S::NO; ++k) A.lo[k] = T[k * STRIDE];
#pragma unroll
    for (int k = 0; k < M * M; ++k) A.di[k] = T[(S::O_DI + k) * STRIDE];
#pragma unroll
    for (int k = 0; k < S::NO; ++k) A.up[k] = T[(S::O_UP + k) * STRIDE];
}

// e = local element index (addresses the element tiles), eg = global element index (addresses the
// pattern table of a translation-invariant level, PatOp)
template <int M, int B, int ST, bool DIAG>
__device__ __forceinline__ void load_blocks(const double* __restrict__ mat, const PatOp& po, int64_t e,
                                            int64_t eg, bool active, RegOp<M, ST>& A, double (*ds)[B],
                                            bool rec = false) {
    using S = OpShape<M, ST>;
    constexpr int ND = DIAG ? M : M * M;
    constexpr int K = S::O_DV + ND;
    const int t = threadIdx.x;
    if (po.tab != nullptr) active = active && eg >= 0 && eg < po.n_glob;
    if (active) {
        if (po.tab != nullptr) {
            const int64_t s = eg < po.n_head ? eg
                            : (eg >= po.n_glob - po.n_tail ? po.n_head + 1 + (eg - (po.n_glob - po.n_tail))
                                                           : po.n_head);
            load_blocks_from<M, B, ST, DIAG, 1>(po.tab + s * K, A, ds);
        } else {
            load_blocks_from<M, B, ST, DIAG, AMG1D_TILE>(mat + (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31), A, ds,
                                                         rec);
        }
    } else {
#pragma unroll
        for (int k = 0; k < ND; ++k) ds[k][t] = 0.0;
#pragma unroll
        for (int k = 0; k < S::NO; ++k) { A.lo[k] = 0.0; A.up[k] = 0.0; }
#pragma unroll
        for (int k = 0; k < M * M; ++k) A.di[k] = 0.0;
    }
}

// Resident CTAs per SM the register allocation aims for.  Dense 4x4 blocks keep 48 operator doubles in
// registers (164 registers, 3 CTAs of 128 threads); the compressed structure classes keep 24 and fit 5
// CTAs (<= 102 registers, no spills), which is what hides the compute phase of one CTA behind the load
// phase of the others (measured on B200, T level 0: 5.0 -> 4.5 ms per leg; 6 CTAs spill and lose).  The
// 2x2 and 1x1 kernels are capped at 64 registers for 8 CTAs/SM.
#ifndef FUSED_MINB_SMALL
#define FUSED_MINB_SMALL 8   // m <= 2: <= 64 registers, no spills; measured best of 3 / 8 / 10 / 12 (T: -4 %)
#endif
constexpr int fused_min_blocks(int m, int st) {
    return m >= 5 ? (st != AMG1D_ST_DENSE ? 3 : 2)
         : m == 4 ? (st != AMG1D_ST_DENSE ? 5 : 3)
         : m <= 2 ? FUSED_MINB_SMALL : FUSED_MINB;
}
#define FUSED_BOUNDS(M) __launch_bounds__(B, fused_min_blocks(M, ST))
// The constant-operand legs (f_down_c / f_up_c) keep no operator in registers on their interior path, so
// they can be given more resident CTAs than f_down / f_up; their table path (the few CTAs at the ends of a
// level) then spills, which costs nothing measurable.  0 = the bounds of f_down / f_up.
#ifndef FUSED_C_MINB3
#define FUSED_C_MINB3 8   // C3 level 1 (3 x 3 dense): 0.535 / 0.546 ms per leg at 3 CTAs/SM, 0.369 / 0.364 at 8
#endif
#ifndef FUSED_C_MINB4
#define FUSED_C_MINB4 8   // measured on B200 (T level 0, profiles/r01e_sweep_const.jsonl): 5 CTAs/SM (the bounds
#endif                    // of f_down / f_up) 2.31 / 2.03 ms per leg, 6: 2.12 / 1.94, 8 (64 registers): 1.99 / 1.92
#ifndef FUSED_C_MINB5
#define FUSED_C_MINB5 0   // 5 x 5: the interior path needs the registers (6 CTAs/SM: 1.86 / 1.81 ms against 1.01 / 0.98;
#endif                    // profiles/r01f_sweep_const.jsonl)
constexpr int fused_c_min_blocks(int m, int st) {
    return (m == 3 && FUSED_C_MINB3) ? FUSED_C_MINB3 : (m == 4 && FUSED_C_MINB4) ? FUSED_C_MINB4
         : (m >= 5 && FUSED_C_MINB5) ? FUSED_C_MINB5 : fused_min_blocks(m, st);
}
#define FUSED_C_BOUNDS(M) __launch_bounds__(B, fused_c_min_blocks(M, ST))

// nsweep pre-smoothing sweeps, residual, restriction to the coarse right-hand side.
//   halo window elements on each side are recomputed (nsweep + 1, plus `ratio` for two-parent
//   transfers whose coarse elements also gather from the children of their left neighbour);
//   out = elements emitted per CTA (a multiple of the agglomeration ratio).
//   Coarse element Kc is gathered by the thread of its first P0-child, in the order of g_restrict:
//   the P1 blocks of the children of Kc - 1, then the P0 blocks of its own children.
// The leg after the operator has been placed (A: RegOp in registers + Dinv column dcol in shared memory, or
// ParamOp in the constant bank); everything from the dependency wait on.
template <int M, int MC, int B, int ST, bool DIAG, class OP>
__device__ __forceinline__ void
down_leg(const OP& A, const double* __restrict__ dcol, Exchange<M, B>& ex, double (*rs)[B + 8], int64_t e,
         bool active, int ilo, int iup, const double* __restrict__ b, const double* __restrict__ xin,
         double* __restrict__ xout, const double* __restrict__ P0, const double* __restrict__ P1,
         const TransferMap& tm, double* __restrict__ rc, int64_t n, double alpha, int nsweep, int zero_guess,
         const WinIdx& wi, const Slab& sl, const HaloLeg& hl) {
    const int t = threadIdx.x;
    const int halo = wi.halo, out = wi.out;
    double bb[M], xc[M], xl[M], xr[M];
    pdl_wait();                                                  // b, x and everything written below are not
    {   // peer-memory exchange: CTAs whose window reaches a slab edge wait for the neighbour's edge (halo_p2p.cuh)
        const int64_t e0 = (int64_t)blockIdx.x * out - halo;
        halo_leg_wait(hl, sl.gl > 0 && e0 < 0, sl.gr > 0 && e0 + B > n);
    }
    if (active) {
        load_vec<M>(b + e * M, bb);
        if (zero_guess) {
#pragma unroll
            for (int i = 0; i < M; ++i) xc[i] = 0.0;
        } else {
            load_vec<M>(xin + e * M, xc);
        }
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) { bb[i] = 0.0; xc[i] = 0.0; }
    }
    cp_async_commit_wait_all();  // this thread's Dinv column has landed (only this thread reads it)
    __syncthreads();             // exch_init visible
    int buf = 0;
    for (int s = 0; s < nsweep; ++s) {
        const bool zg = zero_guess && s == 0;
        if (!zg) {
            exchange<M, B, ST>(ex, buf, ilo, iup, xc, xl, xr);
            buf ^= 1;
        }
        reg_sweep<M, ST, DIAG, B>(A, ilo, iup, dcol, bb, xl, xc, xr, alpha, zg);
    }
    const bool mine = t >= halo && t < halo + out;              // this CTA's share of the level (e >= 0)
    if (mine && e < n) {
        store_vec<M>(xout + e * M, xc);
        if (hl.on) halo_leg_push<M>(xc, e, n, hl.gd, hl.x_left, hl.x_right, hl.f_left, hl.f_right);
    }
    // residual with the final iterate, then restriction
    exchange<M, B, ST>(ex, buf, ilo, iup, xc, xl, xr);
    double r[M];
    reg_residual<M, ST>(A, ilo, iup, bb, xl, xc, xr, r);
#pragma unroll
    for (int i = 0; i < M; ++i) rs[i][t] = r[i];
    __syncthreads();
    const int64_t eg = e + sl.e_off;                            // global element index
    if (mine && eg < tm.n_fine) {
        int kdiv, kmod;                                         // (eg + shift) = ratio * (.. + kdiv) + kmod
        small_divmod(wi.qmod0 + t, tm.ratio, &kdiv, &kmod);
        if (kmod == 0) {
            const int64_t Kc = (int64_t)blockIdx.x * wi.opr + wi.qdiv0 + kdiv + tm.base;   // eg == tm.first(Kc)
            const int64_t Kl = Kc - sl.c_off;                   // local coarse element index
            if (Kc >= 0 && Kc < tm.n_coarse && Kl >= 0 && Kl < sl.nc) {
                double acc[MC];
#pragma unroll
                for (int j = 0; j < MC; ++j) acc[j] = 0.0;
                if (P1) {                                       // children of Kc - 1: [eg - ratio, eg), clamped
                    const int k0 = eg >= tm.ratio ? -tm.ratio : -(int)eg;
                    for (int k = k0; k < 0; ++k) {
                        const double* P = P1 + win_blk(tm, wi, eg + k, t + k) * (M * MC);
                        const int w = t + k;
#pragma unroll
                        for (int j = 0; j < MC; ++j)
#pragma unroll
                            for (int i = 0; i < M; ++i) acc[j] = fma(P[j * M + i], rs[i][w], acc[j]);
                    }
                }
                const int k1 = eg + tm.ratio <= tm.n_fine ? tm.ratio : (int)(tm.n_fine - eg);   // own children
                for (int k = 0; k < k1; ++k) {
                    const double* P = P0 + win_blk(tm, wi, eg + k, t + k) * (M * MC);
                    const int w = t + k;
#pragma unroll
                    for (int j = 0; j < MC; ++j)
#pragma unroll
                        for (int i = 0; i < M; ++i) acc[j] = fma(P[j * M + i], rs[i][w], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < MC; ++j) rc[Kl * MC + j] = acc[j];
                if (hl.on) halo_leg_push<MC>(acc, Kl, hl.nc, hl.gd, hl.c_left, hl.c_right, hl.f_left, hl.f_right);
            }
        }
    }
}

template <int M, int MC, int B, int ST, bool DIAG>
__global__ void FUSED_BOUNDS(M)
f_down(const double* __restrict__ mat, PatOp po, int ilo, int iup, const double* __restrict__ b,
       const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
       const double* __restrict__ P1, TransferMap tm, double* __restrict__ rc, int64_t n, double alpha,
       int nsweep, int zero_guess, WinIdx wi, Slab sl, int rec, const __grid_constant__ HaloLeg hl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double rs[M][B + 8];
    __shared__ double ds[DIAG ? M : M * M][B];
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * wi.out - wi.halo + t;     // local element index (ghosts < 0, >= n)
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    exch_init<M, B>(ex);
    RegOp<M, ST> A;
    load_blocks<M, B, ST, DIAG>(mat, po, e, e + sl.e_off, active, A, ds, rec != 0);   // operator: independent of earlier kernels
    recompute_dinv<M, B, ST, DIAG>(A, ds, active, rec);          // before the dependency wait: few registers are live yet
    down_leg<M, MC, B, ST, DIAG>(A, &ds[0][t], ex, rs, e, active, ilo, iup, b, xin, xout, P0, P1, tm, rc, n, alpha,
                                 nsweep, zero_guess, wi, sl, hl);
}

// CTA-uniform test of the constant-operand legs: the whole window of this CTA is active (inside the slab
// and its ghosts) and lies in the translation-invariant interior of the level, i.e. every thread would
// fetch row n_head of the pattern table.
__device__ __forceinline__ bool window_is_interior(const PatOp& po, const WinIdx& wi, const Slab& sl, int64_t n,
                                                   int B) {
    const int64_t e0 = (int64_t)blockIdx.x * wi.out - wi.halo;        // local element of thread 0
    const int64_t g0 = e0 + sl.e_off;                                 // its global number
    return e0 >= -(int64_t)sl.gl && e0 + B <= n + sl.gr && g0 >= po.n_head && g0 + B <= po.n_glob - po.n_tail;
}

// f_down with the interior block set of a pattern level as a by-value kernel parameter (ParamOp): interior
// CTAs run the leg with constant-bank operands, the few CTAs that touch the head / tail block sets, a slab
// end or the end of the level take the register path of f_down (pattern table, po.tab != nullptr).
template <int M, int MC, int B, int ST, bool DIAG>
__global__ void FUSED_C_BOUNDS(M)
f_down_c(const __grid_constant__ ParamOp<M, ST, DIAG> pk, PatOp po, int ilo, int iup,
         const double* __restrict__ b, const double* __restrict__ xin, double* __restrict__ xout,
         const double* __restrict__ P0, const double* __restrict__ P1, TransferMap tm, double* __restrict__ rc,
         int64_t n, double alpha, int nsweep, int zero_guess, WinIdx wi, Slab sl,
         const __grid_constant__ HaloLeg hl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double rs[M][B + 8];
    __shared__ double ds[DIAG ? M : M * M][B];
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * wi.out - wi.halo + t;
    exch_init<M, B>(ex);
    if (window_is_interior(po, wi, sl, n, B)) {
        down_leg<M, MC, B, ST, DIAG>(pk, nullptr, ex, rs, e, true, ilo, iup, b, xin, xout, P0, P1, tm, rc, n, alpha,
                                     nsweep, zero_guess, wi, sl, hl);
    } else {
        const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
        RegOp<M, ST> A;
        load_blocks<M, B, ST, DIAG>(nullptr, po, e, e + sl.e_off, active, A, ds);
        down_leg<M, MC, B, ST, DIAG>(A, &ds[0][t], ex, rs, e, active, ilo, iup, b, xin, xout, P0, P1, tm, rc, n,
                                     alpha, nsweep, zero_guess, wi, sl, hl);
    }
}

// prolongation + correction, nsweep post-smoothing sweeps, optional || b - A x ||^2 partial sums.
template <int M, int MC, int B, int ST, bool DIAG, class OP>
__device__ __forceinline__ void
up_leg(const OP& A, const double* __restrict__ dcol, Exchange<M, B>& ex, int64_t e, bool active, int ilo, int iup,
       const double* __restrict__ b, const double* __restrict__ xin, double* __restrict__ xout,
       const double* __restrict__ P0, const double* __restrict__ P1, const TransferMap& tm,
       const double* __restrict__ xcoarse, int64_t n, double alpha, int nsweep, const WinIdx& wi,
       double* __restrict__ partial, const Slab& sl, const HaloLeg& hl) {
    const int t = threadIdx.x;
    const int halo = wi.halo, out = wi.out;
    double bb[M], xc[M], xl[M], xr[M];
    pdl_wait();
    {   // peer-memory exchange: the CTAs near a slab edge read fine ghosts and / or ghosts of the coarse correction
        const int64_t e0 = (int64_t)blockIdx.x * out - halo;
        halo_leg_wait(hl, sl.gl > 0 && e0 < tm.ratio + 1, sl.gr > 0 && e0 + B + tm.ratio + 1 > n);
    }
    if (active) {
        load_vec<M>(b + e * M, bb);
        load_vec<M>(xin + e * M, xc);
        // x += P0 x_c[parent] (+ P1 x_c[parent + 1])   (same operation order as g_prolong)
        const int64_t eg = e + sl.e_off;
        const int64_t pb = win_blk(tm, wi, eg, t) * (M * MC);
        int kdiv, kmod;
        small_divmod(wi.qmod0 + t, tm.ratio, &kdiv, &kmod);
        const int64_t par = (int64_t)blockIdx.x * wi.opr + wi.qdiv0 + kdiv + tm.base;   // == tm.par(eg)
        const double* c0 = xcoarse + (par - sl.c_off) * MC;
        double y[M];
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < MC; ++j) {
            const double cj = c0[j];
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(P0[pb + j * M + i], cj, y[i]);
        }
        if (P1) {
#pragma unroll
            for (int j = 0; j < MC; ++j) {
                const double cj = c0[MC + j];
#pragma unroll
                for (int i = 0; i < M; ++i) y[i] = fma(P1[pb + j * M + i], cj, y[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < M; ++i) xc[i] = xc[i] + y[i];
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) { bb[i] = 0.0; xc[i] = 0.0; }
    }
    cp_async_commit_wait_all();
    __syncthreads();
    int buf = 0;
    for (int s = 0; s < nsweep; ++s) {
        exchange<M, B, ST>(ex, buf, ilo, iup, xc, xl, xr);
        buf ^= 1;
        reg_sweep<M, ST, DIAG, B>(A, ilo, iup, dcol, bb, xl, xc, xr, alpha, false);
    }
    const bool emit = e < n && t >= halo && t < halo + out;     // e >= 0 for these threads
    if (emit) {
        store_vec<M>(xout + e * M, xc);
        if (hl.on) halo_leg_push<M>(xc, e, n, hl.gd, hl.x_left, hl.x_right, hl.f_left, hl.f_right);
    }
    if (partial) {
        exchange<M, B, ST>(ex, buf, ilo, iup, xc, xl, xr);
        double r[M];
        reg_residual<M, ST>(A, ilo, iup, bb, xl, xc, xr, r);
        double s2 = 0.0;
        if (emit) {
#pragma unroll
            for (int i = 0; i < M; ++i) s2 = fma(r[i], r[i], s2);
        }
        s2 = block_sum(s2);
        if (t == 0) partial[blockIdx.x] = s2;
    }
}

template <int M, int MC, int B, int ST, bool DIAG>
__global__ void FUSED_BOUNDS(M)
f_up(const double* __restrict__ mat, PatOp po, int ilo, int iup, const double* __restrict__ b,
     const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
     const double* __restrict__ P1, TransferMap tm, const double* __restrict__ xcoarse, int64_t n,
     double alpha, int nsweep, WinIdx wi, double* __restrict__ partial, Slab sl, int rec,
     const __grid_constant__ HaloLeg hl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double ds[DIAG ? M : M * M][B];
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * wi.out - wi.halo + t;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    exch_init<M, B>(ex);
    RegOp<M, ST> A;
    load_blocks<M, B, ST, DIAG>(mat, po, e, e + sl.e_off, active, A, ds, rec != 0);   // operator: independent of earlier kernels
    recompute_dinv<M, B, ST, DIAG>(A, ds, active, rec);          // before the dependency wait: few registers are live yet
    up_leg<M, MC, B, ST, DIAG>(A, &ds[0][t], ex, e, active, ilo, iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha,
                               nsweep, wi, partial, sl, hl);
}

// f_up with constant-bank operands for the interior CTAs of a pattern level (see f_down_c)
template <int M, int MC, int B, int ST, bool DIAG>
__global__ void FUSED_C_BOUNDS(M)
f_up_c(const __grid_constant__ ParamOp<M, ST, DIAG> pk, PatOp po, int ilo, int iup, const double* __restrict__ b,
       const double* __restrict__ xin, double* __restrict__ xout, const double* __restrict__ P0,
       const double* __restrict__ P1, TransferMap tm, const double* __restrict__ xcoarse, int64_t n,
       double alpha, int nsweep, WinIdx wi, double* __restrict__ partial, Slab sl,
       const __grid_constant__ HaloLeg hl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double ds[DIAG ? M : M * M][B];
    pdl_launch_dependents();
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * wi.out - wi.halo + t;
    exch_init<M, B>(ex);
    if (window_is_interior(po, wi, sl, n, B)) {
        up_leg<M, MC, B, ST, DIAG>(pk, nullptr, ex, e, true, ilo, iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha,
                                   nsweep, wi, partial, sl, hl);
    } else {
        const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
        RegOp<M, ST> A;
        load_blocks<M, B, ST, DIAG>(nullptr, po, e, e + sl.e_off, active, A, ds);
        up_leg<M, MC, B, ST, DIAG>(A, &ds[0][t], ex, e, active, ilo, iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha,
                                   nsweep, wi, partial, sl, hl);
    }
}

// residual + restriction, streaming (single-parent transfers; used when the multi-sweep kernel does
// not apply)
template <int M, int MC, int B, int ST>
__global__ void __launch_bounds__(B)
f_residual_restrict(const double* __restrict__ mat, int K, int ilo, int iup, const double* __restrict__ b,
                    const double* __restrict__ x, const double* __restrict__ P0, TransferMap tm,
                    double* __restrict__ rc, int64_t n, int out) {
    __shared__ double rs[M][B + 8];
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * out + t;
    const bool active = e < n && t < out;
    double r[M];
    if (e < n) {
        const double* T = mat + (e >> 5) * (int64_t)K * AMG1D_TILE + (e & 31);
        double xl[M], xc[M], xr[M], y[M], bb[M];
        load_vec<M>(b + e * M, bb);
        load_vec<M>(x + e * M, xc);
        load_neighbours<M, ST>(x, e, ilo, iup, xl, xr);
        stream_Ax<M, ST>(T, ilo, iup, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) rs[i][t] = r[i];
    __syncthreads();
    const int ratio = tm.ratio;
    if (active && (e % ratio) == 0) {
        double acc[MC];
#pragma unroll
        for (int j = 0; j < MC; ++j) acc[j] = 0.0;
        for (int c = 0; c < ratio && e + c < n; ++c) {
            const double* P = P0 + tm.blk(e + c) * (M * MC);
#pragma unroll
            for (int j = 0; j < MC; ++j)
#pragma unroll
                for (int i = 0; i < M; ++i) acc[j] = fma(P[j * M + i], rs[i][t + c], acc[j]);
        }
        const int64_t Kc = e / ratio;
#pragma unroll
        for (int j = 0; j < MC; ++j) rc[Kc * MC + j] = acc[j];
    }
}

// ---- single-CTA coarse tail ----------------------------------------------------------------------------
// The sub-V-cycle of the coarse levels (at most TAIL_B elements each, one block size M, block
// smoothers, single-parent closed-form transfers) in ONE CTA that works entirely out of shared memory:
//   * at entry every tail level's operator tiles (and transfer blocks) are copied global -> shared with
//     cp.async - BEFORE the programmatic-dependency wait, so under PDL this prefetch overlaps the last
//     f_down of the levels above;
//   * one thread per element; per leg the element's blocks go shared -> registers, iterates are
//     exchanged through shared memory, the right-hand sides and pre-smoothed iterates of all tail
//     levels stay in shared memory between the down and the up leg;
//   * the only dependent global traffic is the first level's right-hand side in and its iterate out.
// It replaces two launches per level (each a few microseconds of latency for a few hundred bytes of
// work) by one launch.  Arithmetic per element is that of f_down / f_up / g_coarse_solve, so the
// results are bit-identical to the per-level kernels.
#define TAIL_MAXL 16
struct TailLevel {
    MatDesc md;
    const double* mat;
    int64_t n;
    double* x;          // iterate buffer 0 of the level (element 0); written for the first tail level only
    double* b;          // right-hand side; read for the first tail level only
    TransferMap tm;     // transfer to the next tail level (unused on the last one)
    const double* P0;
    int op_off, op_len; // this level's tiles inside the shared operator area (doubles)
    int vec_off;        // this level's vectors inside the shared b / x areas (doubles)
    int p_off, p_len;   // this level's transfer blocks inside the shared P area (doubles)
};

// shared-memory carve-up of f_tail (in doubles, after the descriptors)
struct TailPlan {
    int n_desc, ex, rs, cw, ops, bvec, xvec, pblk, total;   // offsets; total = size in doubles
};
inline TailPlan tail_plan(int m, int op_total, int vec_total, int p_total) {
    TailPlan p;
    int o = 0;
    p.n_desc = o; o += (int)((TAIL_MAXL * sizeof(TailLevel) + 7) / 8);
    p.ex = o; o += 2 * m * (TAIL_B + 2);
    p.rs = o; o += m * (TAIL_B + 8);
    p.cw = o; o += 64;
    o = (o + 1) & ~1;                                   // 16-byte alignment for cp.async
    p.ops = o; o += op_total;
    p.pblk = o; o += (p_total + 1) & ~1;
    p.bvec = o; o += vec_total;
    p.xvec = o; o += vec_total;
    p.total = o;
    return p;
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// Dense register copy of the element's operator whatever the level's structure class (entries that
// the class does not store are exact zeros, so the dense chains give identical results).  T points at
// the element's lane of its (shared-memory) tile.
template <int M>
__device__ __forceinline__ void tail_load(const MatDesc& d, const double* T, bool active, RegOp<M, 0>& A) {
    if (active) {
#pragma unroll
        for (int k = 0; k < M * M; ++k) A.di[k] = T[(d.o_di + k) * AMG1D_TILE];
        if (d.st == AMG1D_ST_DENSE) {
#pragma unroll
            for (int k = 0; k < M * M; ++k) { A.lo[k] = T[k * AMG1D_TILE]; A.up[k] = T[(d.o_up + k) * AMG1D_TILE]; }
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if (d.st == AMG1D_ST_COLROW) {
                        A.lo[j * M + i] = (j == d.ilo) ? T[i * AMG1D_TILE] : 0.0;
                        A.up[j * M + i] = (i == d.iup) ? T[(d.o_up + j) * AMG1D_TILE] : 0.0;
                    } else {
                        A.lo[j * M + i] = (i == d.ilo) ? T[j * AMG1D_TILE] : 0.0;
                        A.up[j * M + i] = (j == d.iup) ? T[(d.o_up + i) * AMG1D_TILE] : 0.0;
                    }
                }
        }
    } else {
#pragma unroll
        for (int k = 0; k < M * M; ++k) { A.lo[k] = 0.0; A.di[k] = 0.0; A.up[k] = 0.0; }
    }
}

// one damped block-Jacobi sweep with Dinv read from the element's shared tile (inactive threads: x stays 0)
template <int M>
__device__ __forceinline__ void tail_sweep(const RegOp<M, 0>& A, const double* dcol, bool active,
                                           const double (&bb)[M], const double (&xl)[M], double (&xc)[M],
                                           const double (&xr)[M], double alpha, bool zero_guess) {
    if (active) reg_sweep<M, 0, false, AMG1D_TILE>(A, 0, 0, dcol, bb, xl, xc, xr, alpha, zero_guess);
}

template <int M>
__global__ void __launch_bounds__(TAIL_B, 1)
f_tail(const TailLevel* __restrict__ lv, int nl, const double* __restrict__ fac, int nPre, int nPost,
       double alpha, TailPlan pl) {
    extern __shared__ __align__(16) double tsm[];
    pdl_launch_dependents();
    const int t = threadIdx.x;
    TailLevel* sd = reinterpret_cast<TailLevel*>(tsm + pl.n_desc);
    Exchange<M, TAIL_B>& ex = *reinterpret_cast<Exchange<M, TAIL_B>*>(tsm + pl.ex);
    double (*rs)[TAIL_B + 8] = reinterpret_cast<double (*)[TAIL_B + 8]>(tsm + pl.rs);
    double* cw = tsm + pl.cw;
    double* ops = tsm + pl.ops;
    double* pblk = tsm + pl.pblk;
    double* bvec = tsm + pl.bvec;
    double* xvec = tsm + pl.xvec;
    // descriptors, then every level's operator tiles and transfer blocks: read-only data, fetched before
    // the dependency wait
    for (int i = t; i < (int)(nl * sizeof(TailLevel) / 8); i += TAIL_B)
        (tsm + pl.n_desc)[i] = reinterpret_cast<const double*>(lv)[i];
    exch_init<M, TAIL_B>(ex);
    __syncthreads();
    for (int l = 0; l < nl; ++l) {
        const TailLevel& L = sd[l];
        for (int i = 2 * t; i < L.op_len; i += 2 * TAIL_B) cp_async16(ops + L.op_off + i, L.mat + i);
        for (int i = t; i < L.p_len; i += TAIL_B) cp_async8(pblk + L.p_off + i, L.P0 + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pdl_wait();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    RegOp<M, 0> A;
    double bb[M], xc[M], xl[M], xr[M];
    int buf = 0;
    // ---- down: zero guess, nPre sweeps, residual, restriction ----
    for (int l = 0; l < nl - 1; ++l) {
        const TailLevel& L = sd[l];
        const bool active = t < L.n;
        const double* T = ops + L.op_off + (t >> 5) * (L.md.K * AMG1D_TILE) + (t & 31);
        const double* dcol = T + L.md.o_dv * AMG1D_TILE;
        double* bl = bvec + L.vec_off;
        tail_load<M>(L.md, T, active, A);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            bb[i] = active ? (l == 0 ? L.b[t * M + i] : bl[t * M + i]) : 0.0;
            xc[i] = 0.0;
        }
        if (l == 0 && active) {
#pragma unroll
            for (int i = 0; i < M; ++i) bl[t * M + i] = bb[i];       // kept for the up leg
        }
        for (int s = 0; s < nPre; ++s) {
            if (s > 0) {
                exchange<M, TAIL_B, 0>(ex, buf, 0, 0, xc, xl, xr);
                buf ^= 1;
            }
            tail_sweep<M>(A, dcol, active, bb, xl, xc, xr, alpha, s == 0);
        }
        if (active) {
#pragma unroll
            for (int i = 0; i < M; ++i) xvec[L.vec_off + t * M + i] = xc[i];
        }
        exchange<M, TAIL_B, 0>(ex, buf, 0, 0, xc, xl, xr);
        buf ^= 1;
        double r[M];
        reg_residual<M, 0>(A, 0, 0, bb, xl, xc, xr, r);
#pragma unroll
        for (int i = 0; i < M; ++i) rs[i][t] = r[i];
        __syncthreads();
        const int ratio = L.tm.ratio;
        if (active && (t % ratio) == 0) {
            double acc[M];
#pragma unroll
            for (int j = 0; j < M; ++j) acc[j] = 0.0;
            for (int c = 0; c < ratio && t + c < L.n; ++c) {
                const double* P = pblk + L.p_off + L.tm.blk(t + c) * (M * M);
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int i = 0; i < M; ++i) acc[j] = fma(P[j * M + i], rs[i][t + c], acc[j]);
            }
            double* bc = bvec + sd[l + 1].vec_off + (t / ratio) * M;
#pragma unroll
            for (int j = 0; j < M; ++j) bc[j] = acc[j];
        }
        __syncthreads();
    }
    // ---- coarsest level: block-Thomas substitution by one warp ----
    {
        const TailLevel& L = sd[nl - 1];
        if (t < 32) coarse_solve_warp(fac, M, L.n, bvec + L.vec_off, xvec + L.vec_off, cw, cw + 32);
        __syncthreads();
    }
    // ---- up: prolongation + correction, nPost sweeps ----
    for (int l = nl - 2; l >= 0; --l) {
        const TailLevel& L = sd[l];
        const bool active = t < L.n;
        const double* T = ops + L.op_off + (t >> 5) * (L.md.K * AMG1D_TILE) + (t & 31);
        const double* dcol = T + L.md.o_dv * AMG1D_TILE;
        tail_load<M>(L.md, T, active, A);
        if (active) {
            const double* P = pblk + L.p_off + L.tm.blk(t) * (M * M);
            const double* c0 = xvec + sd[l + 1].vec_off + (t / L.tm.ratio) * M;
            double y[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                bb[i] = bvec[L.vec_off + t * M + i];
                xc[i] = xvec[L.vec_off + t * M + i];
                y[i] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const double cj = c0[j];
#pragma unroll
                for (int i = 0; i < M; ++i) y[i] = fma(P[j * M + i], cj, y[i]);
            }
#pragma unroll
            for (int i = 0; i < M; ++i) xc[i] = xc[i] + y[i];
        } else {
#pragma unroll
            for (int i = 0; i < M; ++i) { bb[i] = 0.0; xc[i] = 0.0; }
        }
        for (int s = 0; s < nPost; ++s) {
            exchange<M, TAIL_B, 0>(ex, buf, 0, 0, xc, xl, xr);
            buf ^= 1;
            tail_sweep<M>(A, dcol, active, bb, xl, xc, xr, alpha, false);
        }
        if (active) {
#pragma unroll
            for (int i = 0; i < M; ++i) xvec[L.vec_off + t * M + i] = xc[i];
            if (l == 0) store_vec<M>(L.x + t * M, xc);
        }
        __syncthreads();
    }
}

// ---- host-side dispatch ---------------------------------------------------------------------------------
#define FUSED_FOR_M(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)

// structure classes need m >= 2, and ST_COLROW the DG trace row (see FUSED_COLROW_IUP)
inline bool fast_tier_ok(const MatDesc& d) {
    if (d.st == AMG1D_ST_DENSE) return true;
    if (d.m < 2) return false;
    return d.st != AMG1D_ST_COLROW || d.iup == FUSED_COLROW_IUP;
}

inline bool fused_sweep(const MatDesc& d, const double* mat, const double* b, const double* xin,
                        double* xout, int64_t n, double alpha, int zero_guess, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (!fast_tier_ok(d)) return false;
    switch ((d.m * 2 + d.diag) * 4 + d.st) {
#define Y(MM, DG, SS)                                                                                     \
    case (MM * 2 + DG) * 4 + SS:                                                                          \
        f_sweep<MM, DG != 0, SS><<<grid, 256, 0, st>>>(mat, d.ilo, d.iup, b, xin, xout, n, alpha, zero_guess); \
        return true;
#define X(MM) Y(MM, 0, 0) Y(MM, 1, 0) Y(MM, 0, 1) Y(MM, 1, 1) Y(MM, 0, 2) Y(MM, 1, 2)
        FUSED_FOR_M(X)
#undef X
#undef Y
        default: return false;
    }
}

inline bool fused_resnorm(const MatDesc& d, const double* mat, const double* b, const double* x,
                          int64_t n, double* partial, int64_t partial_cap, int* nblocks,
                          cudaStream_t st) {
    const int64_t grid = (n + 255) / 256;
    if (grid > partial_cap || !fast_tier_ok(d)) return false;
    *nblocks = (int)grid;
    switch (d.m * 4 + d.st) {
#define Y(MM, SS)                                                                                       \
    case MM * 4 + SS:                                                                                   \
        f_resnorm<MM, SS><<<(unsigned)grid, 256, 0, st>>>(mat, d.K, d.ilo, d.iup, b, x, n, partial);    \
        return true;
#define X(MM) Y(MM, 0) Y(MM, 1) Y(MM, 2)
        FUSED_FOR_M(X)
#undef X
#undef Y
        default: return false;
    }
}

inline bool fused_matvec_dot(const MatDesc& d, const double* mat, const double* x, double* y, int64_t n,
                             double* partial, int64_t partial_cap, int* nblocks, cudaStream_t st) {
    const int64_t grid = (n + 255) / 256;
    if (grid > partial_cap || !fast_tier_ok(d)) return false;
    *nblocks = (int)grid;
    switch (d.m * 4 + d.st) {
#define Y(MM, SS)                                                                                       \
    case MM * 4 + SS:                                                                                   \
        f_matvec_dot<MM, SS><<<(unsigned)grid, 256, 0, st>>>(mat, d.K, d.ilo, d.iup, x, y, n, partial); \
        return true;
#define X(MM) Y(MM, 0) Y(MM, 1) Y(MM, 2)
        FUSED_FOR_M(X)
#undef X
#undef Y
        default: return false;
    }
}

// (M, MC, ST, DIAG) combinations with a register-resident multi-sweep kernel:
//   block Jacobi, dense and DG-assembled structure (DG / agglomerated levels);
//   point Jacobi on CG levels in group form (m = 1: CG p=1; m >= 2: ST_ROWCOL).
#define FUSED_COMBOS(X)                                                                              \
    X(1, 1, 0, false) X(2, 1, 0, false) X(2, 2, 0, false) X(3, 1, 0, false) X(3, 2, 0, false)         \
    X(4, 2, 0, false) X(4, 3, 0, false) X(5, 3, 0, false)                                             \
    X(2, 1, 1, false) X(2, 2, 1, false) X(3, 1, 1, false) X(3, 2, 1, false)                           \
    X(4, 2, 1, false) X(4, 3, 1, false) X(5, 3, 1, false)                                             \
    X(1, 1, 0, true) X(1, 2, 0, true) X(2, 1, 2, true) X(3, 1, 2, true) X(4, 2, 2, true)

inline int fused_key(int m, int mc, int st, int diag) { return ((m * 16 + mc) * 4 + st) * 2 + (diag ? 1 : 0); }

// halo, elements emitted per CTA and the 32-bit index constants for a leg of nsweep sweeps
inline int64_t floordiv64(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
inline WinIdx fused_window(int nsweep, const TransferMap& tm, bool wide, const Slab& sl, int window = FUSED_B) {
    WinIdx w;
    w.halo = nsweep + 1 + (wide ? tm.ratio : 0);
    w.out = ((window - 2 * w.halo) / tm.ratio) * tm.ratio;
    if (w.out < 1) w.out = 0;
    w.opr = w.out / tm.ratio;
    const int64_t q0 = sl.e_off + tm.shift - w.halo;
    w.qdiv0 = floordiv64(q0, tm.ratio);
    w.qmod0 = (int)(q0 - w.qdiv0 * tm.ratio);
    w.pmod0 = -1;
    if (tm.period > 0 && w.out % tm.period == 0) {
        const int64_t p0 = sl.e_off - w.halo - tm.n_head;
        const int64_t pm = p0 - floordiv64(p0, tm.period) * tm.period;            // floormod
        w.pmod0 = (int)pm + ((tm.ratio + tm.period - 1) / tm.period) * tm.period; // bias: thread offsets >= -ratio
    }
    return w;
}

// tm must be a closed-form map (parent == cp == nullptr); P1 != nullptr selects the two-parent form.
// n_cover: local elements whose coarse parents this rank may have to gather (n, or n + ratio when the
// right slab neighbour's first children contribute to this rank's last coarse element).
// Return value of fused_down / fused_up: FUSED_NA = no such kernel (the caller takes the streaming /
// generic path), FUSED_OK = launched, FUSED_ERR = the launch itself failed (*err holds the CUDA error;
// never silently papered over by a fall-back).
enum { FUSED_NA = 0, FUSED_OK = 1, FUSED_ERR = -1 };

inline int fused_down(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, bool zero,
                       const double* mat, const PatOp& po, const double* b, const double* xin, double* xout,
                       const double* P0, const double* P1, double* rc, int64_t n, int64_t n_cover,
                       double alpha, const Slab& sl, cudaStream_t st, bool pdl, cudaError_t* err, int rec = 0,
                       const HaloLeg& hl = HaloLeg()) {
    const WinIdx w = fused_window(nsweep, tm, P1 != nullptr || tm.shift != 0 || tm.base != 0, sl);
    if (w.out < tm.ratio || w.out < FUSED_B / 2 || !fast_tier_ok(d)) return FUSED_NA;
    const unsigned grid = (unsigned)((n_cover + w.out - 1) / w.out);
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0):                                                  \
        if (po.tab && po.host_interior && sizeof(ParamOp<MM, SS, DG>) == (size_t)d.K * 8) {                                                                \
            ParamOp<MM, SS, DG> pk;                                                                      \
            memcpy(&pk, po.host_interior, sizeof(pk));                                                   \
            *err = launch_fused(f_down_c<MM, MCC, FUSED_B, SS, DG>, grid, FUSED_B, 0, st, pdl, pk, po, d.ilo, \
                                d.iup, b, xin, xout, P0, P1, tm, rc, n, alpha, nsweep, zero ? 1 : 0, w, sl, hl); \
        } else                                                                                           \
        *err = launch_fused(f_down<MM, MCC, FUSED_B, SS, DG>, grid, FUSED_B, 0, st, pdl, mat, po, d.ilo, d.iup, \
                            b, xin, xout, P0, P1, tm, rc, n, alpha, nsweep, zero ? 1 : 0, w, sl,         \
                            (!po.tab && !DG) ? rec : 0, hl);                                             \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;
        FUSED_COMBOS(X)
#undef X
        default: return FUSED_NA;
    }
}

inline int fused_up(const MatDesc& d, int mc, const TransferMap& tm, int nsweep, const double* mat,
                     const PatOp& po, const double* b, const double* xin, double* xout, const double* P0,
                     const double* P1, const double* xcoarse, int64_t n, double alpha, double* partial,
                     int64_t partial_cap, int* nblocks, const Slab& sl, cudaStream_t st, bool pdl,
                     cudaError_t* err, int rec = 0, const HaloLeg& hl = HaloLeg()) {
    const WinIdx w = fused_window(nsweep, tm, false, sl);
    if (w.out < tm.ratio || w.out < FUSED_B / 2 || !fast_tier_ok(d)) return FUSED_NA;
    const int64_t grid = (n + w.out - 1) / w.out;
    if (partial && grid > partial_cap) return FUSED_NA;
    if (nblocks) *nblocks = (int)grid;
    switch (fused_key(d.m, mc, d.st, d.diag)) {
#define X(MM, MCC, SS, DG)                                                                               \
    case ((MM * 16 + MCC) * 4 + SS) * 2 + (DG ? 1 : 0):                                                  \
        if (po.tab && po.host_interior && sizeof(ParamOp<MM, SS, DG>) == (size_t)d.K * 8) {                                                                \
            ParamOp<MM, SS, DG> pk;                                                                      \
            memcpy(&pk, po.host_interior, sizeof(pk));                                                   \
            *err = launch_fused(f_up_c<MM, MCC, FUSED_B, SS, DG>, (unsigned)grid, FUSED_B, 0, st, pdl, pk, po, \
                                d.ilo, d.iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha, nsweep, w, partial, sl, hl); \
        } else                                                                                           \
        *err = launch_fused(f_up<MM, MCC, FUSED_B, SS, DG>, (unsigned)grid, FUSED_B, 0, st, pdl, mat, po,    \
                            d.ilo, d.iup, b, xin, xout, P0, P1, tm, xcoarse, n, alpha, nsweep, w, partial, sl, \
                            (!po.tab && !DG) ? rec : 0, hl);                                             \
        return *err == cudaSuccess ? FUSED_OK : FUSED_ERR;
        FUSED_COMBOS(X)
#undef X
        default: return FUSED_NA;
    }
}

// (M, MC) pairs of the streaming residual + restriction kernel (single-parent transfers)
#define FRR_PAIRS(X) X(1, 1) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(4, 2) X(4, 3) X(5, 3) X(9, 5) X(5, 2) X(9, 2)

inline bool fused_residual_restrict(const MatDesc& d, int mc, const TransferMap& tm, const double* mat,
                                    const double* b, const double* x, const double* P0, double* rc,
                                    int64_t n, cudaStream_t st) {
    const int out = (FUSED_B / tm.ratio) * tm.ratio;
    if (out < tm.ratio || !fast_tier_ok(d)) return false;
    const unsigned grid = (unsigned)((n + out - 1) / out);
    switch ((d.m * 16 + mc) * 4 + d.st) {
#define Y(MM, MCC, SS)                                                                                  \
    case (MM * 16 + MCC) * 4 + SS:                                                                      \
        f_residual_restrict<MM, MCC, FUSED_B, SS><<<grid, FUSED_B, 0, st>>>(mat, d.K, d.ilo, d.iup, b, x, P0, \
                                                                             tm, rc, n, out);           \
        return true;
#define X(MM, MCC) Y(MM, MCC, 0) Y(MM, MCC, 1) Y(MM, MCC, 2)
        FRR_PAIRS(X)
#undef X
#undef Y
        default: return false;
    }
}

// Block sizes with a single-CTA tail kernel.
inline bool tail_supported(int m) { return m == 1 || m == 2; }
#define TAIL_SMEM_MAX (227 * 1024)

template <int M>
inline cudaError_t tail_configure_t(size_t smem) {
    return cudaFuncSetAttribute(f_tail<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// once per device context and tail shape, before the first launch (amg1d_finalize)
inline cudaError_t tail_configure(int m, size_t smem) {
    switch (m) {
        case 1: return tail_configure_t<1>(smem);
        case 2: return tail_configure_t<2>(smem);
        default: return cudaErrorInvalidValue;
    }
}

inline cudaError_t tail_launch(int m, const TailLevel* lv, int nl, const double* fac, int nPre, int nPost,
                               double alpha, const TailPlan& pl, cudaStream_t st, bool pdl) {
    const size_t smem = (size_t)pl.total * 8;
    switch (m) {
        case 1: return launch_fused(f_tail<1>, 1, TAIL_B, smem, st, pdl, lv, nl, fac, nPre, nPost, alpha, pl);
        case 2: return launch_fused(f_tail<2>, 1, TAIL_B, smem, st, pdl, lv, nl, fac, nPre, nPost, alpha, pl);
        default: return cudaErrorInvalidValue;
    }
}
