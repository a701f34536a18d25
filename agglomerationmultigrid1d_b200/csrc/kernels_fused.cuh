// Fast kernel tier: templated on the block size M, one thread per element, operator read straight
// from the element-tile layout with fully coalesced 256-byte warp requests.
//
//  f_sweep / f_resnorm / f_residual_restrict  - streaming kernels (one pass over the operator each)
//  f_down   nPre sweeps + residual + restriction  in ONE pass over the level's operator
//  f_up     prolongation + correction + nPost sweeps (+ optional ||b - A x||^2)  in ONE pass
//
// f_down / f_up keep the element's four M x M blocks (A_lo, A_di, A_up, Dinv) in registers for the
// whole leg and exchange only the M iterate values with the two neighbour threads through shared
// memory between sweeps (Jacobi needs the *old* neighbour values, so the exchange is double
// buffered).  A CTA owns a window of B consecutive elements; after s sweeps only the inner
// [s, B - s) elements are still exact, so the CTA emits the inner B - 2(S+1) elements and adjacent
// CTAs overlap by the halo (the overlap is re-read through L2, not HBM).  Per element the arithmetic
// and its order are identical to the generic tier, so both tiers agree bit for bit on x.
//
// Algorithmic bytes per element (FP64, M x M blocks, coarse block size MC, ratio R children):
//   f_sweep               8 (4 M^2 + 3 M)                         [zero guess: 8 (M^2 + 2 M)]
//   f_down (S sweeps)     8 (4 M^2 + 3 M + MC / R)  (+ P block if it is per-element)
//   f_up   (S sweeps)     8 (4 M^2 + 3 M + MC / R)  (+ P block if it is per-element)
// against S * 8 (4 M^2 + 3 M) + 8 (3 M^2 + 2 M + ...) for the unfused sequence (SURVEY 8d B_ref).
#pragma once
#include <cuda_runtime.h>
#include "kernels_generic.cuh"
#include "layout.cuh"

// Position of a rank's slab inside its level (single GPU: all zero).  Local element index e runs over
// [-gl, n + gr): gl / gr ghost elements (with operator blocks, rhs and iterate) on the left / right
// slab edge; e_off = global index of local element 0; c_off = global index of local coarse element 0.
struct Slab {
    int gl, gr;
    int64_t e_off, c_off;
};

#ifndef FUSED_B
#define FUSED_B 128  // window (threads) per CTA of f_down / f_up
#endif
#ifndef FUSED_MINB
#define FUSED_MINB 3   // __launch_bounds__ min CTAs per SM for f_down / f_up
#endif

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

// y = A_lo xl + A_di xc + A_up xr, operator streamed from the tile (T points at [tile][0][lane]).
template <int M>
__device__ __forceinline__ void stream_Ax(const double* __restrict__ T, const double (&xl)[M],
                                          const double (&xc)[M], const double (&xr)[M],
                                          double (&y)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[(j * M + i) * AMG1D_TILE], xl[j], y[i]);
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[(M * M + j * M + i) * AMG1D_TILE], xc[j], y[i]);
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(T[(2 * M * M + j * M + i) * AMG1D_TILE], xr[j], y[i]);
}

// ---- streaming kernels --------------------------------------------------------------------------------
template <int M, bool DIAG>
__global__ void __launch_bounds__(256) f_sweep(const double* __restrict__ mat,
                                               const double* __restrict__ b,
                                               const double* __restrict__ xin,
                                               double* __restrict__ xout, int64_t n, double alpha,
                                               int zero_guess) {
    constexpr int K = 3 * M * M + (DIAG ? M : M * M);
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
        load_vec<M>(xin + (e - 1) * M, xl);
        load_vec<M>(xin + e * M, xc);
        load_vec<M>(xin + (e + 1) * M, xr);
        stream_Ax<M>(T, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
    }
    double xn[M];
    if constexpr (DIAG) {
#pragma unroll
        for (int i = 0; i < M; ++i)
            xn[i] = __dadd_rn(xc[i], __dmul_rn(alpha, T[(3 * M * M + i) * AMG1D_TILE] * r[i]));
    } else {
        double z[M];
#pragma unroll
        for (int i = 0; i < M; ++i) z[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) z[i] = fma(T[(3 * M * M + j * M + i) * AMG1D_TILE], r[j], z[i]);
#pragma unroll
        for (int i = 0; i < M; ++i) xn[i] = __dadd_rn(xc[i], __dmul_rn(alpha, z[i]));
    }
    store_vec<M>(xout + e * M, xn);
}

// partial[blockIdx] = sum over the block's elements of || b - A x ||^2
template <int M, int KK>
__global__ void __launch_bounds__(256) f_resnorm(const double* __restrict__ mat,
                                                 const double* __restrict__ b,
                                                 const double* __restrict__ x, int64_t n,
                                                 double* __restrict__ partial) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (e < n) {
        const double* T = mat + (e >> 5) * (int64_t)(KK * AMG1D_TILE) + (e & 31);
        double xl[M], xc[M], xr[M], y[M], bb[M];
        load_vec<M>(b + e * M, bb);
        load_vec<M>(x + (e - 1) * M, xl);
        load_vec<M>(x + e * M, xc);
        load_vec<M>(x + (e + 1) * M, xr);
        stream_Ax<M>(T, xl, xc, xr, y);
#pragma unroll
        for (int i = 0; i < M; ++i) { const double r = bb[i] - y[i]; s = fma(r, r, s); }
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
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

// one damped block-Jacobi sweep on register-resident blocks; same operation order as g_sweep
// Dinv is read from shared memory: dcol points at this thread's column of ds[k][thread], stride DS.
template <int M, int DS>
__device__ __forceinline__ void reg_sweep(const double (&A)[3 * M * M], const double* __restrict__ dcol,
                                          const double (&bb)[M], const double (&xl)[M],
                                          double (&xc)[M], const double (&xr)[M], double alpha,
                                          bool zero_guess) {
    double r[M];
    if (zero_guess) {
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - 0.0;
    } else {
        double y[M];
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(A[j * M + i], xl[j], y[i]);
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(A[M * M + j * M + i], xc[j], y[i]);
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
            for (int i = 0; i < M; ++i) y[i] = fma(A[2 * M * M + j * M + i], xr[j], y[i]);
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
    }
    double z[M];
#pragma unroll
    for (int i = 0; i < M; ++i) z[i] = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) z[i] = fma(dcol[(j * M + i) * DS], r[j], z[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) xc[i] = __dadd_rn(xc[i], __dmul_rn(alpha, z[i]));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int M>
__device__ __forceinline__ void reg_residual(const double (&A)[3 * M * M], const double (&bb)[M],
                                             const double (&xl)[M], const double (&xc)[M],
                                             const double (&xr)[M], double (&r)[M]) {
    double y[M];
#pragma unroll
    for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A[j * M + i], xl[j], y[i]);
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A[M * M + j * M + i], xc[j], y[i]);
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = fma(A[2 * M * M + j * M + i], xr[j], y[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = bb[i] - y[i];
}

// publish x into exchange buffer `buf`, barrier, fetch both neighbours
template <int M, int B>
__device__ __forceinline__ void exchange(Exchange<M, B>& ex, int buf, const double (&xc)[M],
                                         double (&xl)[M], double (&xr)[M]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < M; ++i) ex.xs[buf][i][t + 1] = xc[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < M; ++i) {
        xl[i] = ex.xs[buf][i][t];
        xr[i] = ex.xs[buf][i][t + 2];
    }
}

// A_lo / A_di / A_up go to registers; Dinv (used once per sweep) is copied global -> shared with
// cp.async, i.e. without register staging, into this thread's own column ds[k][thread].
template <int M, int B>
__device__ __forceinline__ void load_blocks(const double* __restrict__ mat, int64_t e, bool active,
                                            double (&A)[3 * M * M], double (*ds)[B]) {
    constexpr int K = 4 * M * M;
    const int t = threadIdx.x;
    if (active) {
        const double* T = mat + (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31);
#pragma unroll
        for (int k = 0; k < M * M; ++k) cp_async8(&ds[k][t], T + (3 * M * M + k) * AMG1D_TILE);
#pragma unroll
        for (int k = 0; k < 3 * M * M; ++k) A[k] = T[k * AMG1D_TILE];
    } else {
#pragma unroll
        for (int k = 0; k < M * M; ++k) ds[k][t] = 0.0;
#pragma unroll
        for (int k = 0; k < 3 * M * M; ++k) A[k] = 0.0;
    }
}

// nsweep pre-smoothing sweeps, residual, restriction to the coarse right-hand side.
//   halo = nsweep + 1 window elements on each side are recomputed; out = elements emitted per CTA
//   (a multiple of the agglomeration ratio).
template <int M, int MC, int B>
__global__ void __launch_bounds__(B, (M >= 5 ? 2 : FUSED_MINB))
f_down(const double* __restrict__ mat, const double* __restrict__ b, const double* __restrict__ xin,
       double* __restrict__ xout, const double* __restrict__ P0, TransferMap tm,
       double* __restrict__ rc, int64_t n, double alpha, int nsweep, int zero_guess, int out,
       Slab sl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double rs[M][B + 8];
    __shared__ double ds[M * M][B];
    const int t = threadIdx.x;
    const int halo = nsweep + 1;
    const int64_t e = (int64_t)blockIdx.x * out - halo + t;     // local element index (ghosts < 0, >= n)
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    exch_init<M, B>(ex);
    double A[3 * M * M], bb[M], xc[M], xl[M], xr[M];
    load_blocks<M, B>(mat, e, active, A, ds);
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
            exchange<M, B>(ex, buf, xc, xl, xr);
            buf ^= 1;
        }
        reg_sweep<M, B>(A, &ds[0][t], bb, xl, xc, xr, alpha, zg);
    }
    const bool emit = e >= 0 && e < n && t >= halo && t < halo + out;
    if (emit) store_vec<M>(xout + e * M, xc);
    // residual with the final iterate, then restriction
    exchange<M, B>(ex, buf, xc, xl, xr);
    double r[M];
    reg_residual<M>(A, bb, xl, xc, xr, r);
#pragma unroll
    for (int i = 0; i < M; ++i) rs[i][t] = r[i];
    __syncthreads();
    const int ratio = tm.ratio;
    const int64_t eg = e + sl.e_off;                           // global element index
    if (emit && (eg % ratio) == 0) {
        double acc[MC];
#pragma unroll
        for (int j = 0; j < MC; ++j) acc[j] = 0.0;
        for (int c = 0; c < ratio && e + c < n; ++c) {
            const double* P = P0 + tm.blk(eg + c) * (M * MC);
#pragma unroll
            for (int j = 0; j < MC; ++j)
#pragma unroll
                for (int i = 0; i < M; ++i) acc[j] = fma(P[j * M + i], rs[i][t + c], acc[j]);
        }
        const int64_t Kc = eg / ratio - sl.c_off;              // local coarse element index
#pragma unroll
        for (int j = 0; j < MC; ++j) rc[Kc * MC + j] = acc[j];
    }
}

// prolongation + correction, nsweep post-smoothing sweeps, optional || b - A x ||^2 partial sums.
template <int M, int MC, int B>
__global__ void __launch_bounds__(B, (M >= 5 ? 2 : FUSED_MINB))
f_up(const double* __restrict__ mat, const double* __restrict__ b, const double* __restrict__ xin,
     double* __restrict__ xout, const double* __restrict__ P0, TransferMap tm,
     const double* __restrict__ xcoarse, int64_t n, double alpha, int nsweep, int out,
     double* __restrict__ partial, Slab sl) {
    __shared__ Exchange<M, B> ex;
    __shared__ double ds[M * M][B];
    const int t = threadIdx.x;
    const int halo = nsweep + 1;
    const int64_t e = (int64_t)blockIdx.x * out - halo + t;
    const bool active = e >= -(int64_t)sl.gl && e < n + sl.gr;
    exch_init<M, B>(ex);
    double A[3 * M * M], bb[M], xc[M], xl[M], xr[M];
    load_blocks<M, B>(mat, e, active, A, ds);
    if (active) {
        load_vec<M>(b + e * M, bb);
        load_vec<M>(xin + e * M, xc);
        // x += P x_c   (same operation order as g_prolong: y = sum_j P(i,j) xc_j, then x + y)
        const int64_t eg = e + sl.e_off;
        const double* P = P0 + tm.blk(eg) * (M * MC);
        const double* c0 = xcoarse + (eg / tm.ratio - sl.c_off) * MC;
        double y[M];
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < MC; ++j) {
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
    cp_async_commit_wait_all();
    __syncthreads();
    int buf = 0;
    for (int s = 0; s < nsweep; ++s) {
        exchange<M, B>(ex, buf, xc, xl, xr);
        buf ^= 1;
        reg_sweep<M, B>(A, &ds[0][t], bb, xl, xc, xr, alpha, false);
    }
    const bool emit = e >= 0 && e < n && t >= halo && t < halo + out;
    if (emit) store_vec<M>(xout + e * M, xc);
    if (partial) {
        exchange<M, B>(ex, buf, xc, xl, xr);
        double r[M];
        reg_residual<M>(A, bb, xl, xc, xr, r);
        double s2 = 0.0;
        if (emit) {
#pragma unroll
            for (int i = 0; i < M; ++i) s2 = fma(r[i], r[i], s2);
        }
        s2 = block_sum(s2);
        if (t == 0) partial[blockIdx.x] = s2;
    }
}

// residual + restriction, streaming (used when the multi-sweep kernel does not apply)
template <int M, int MC, int B>
__global__ void __launch_bounds__(B)
f_residual_restrict(const double* __restrict__ mat, int K, const double* __restrict__ b,
                    const double* __restrict__ x, const double* __restrict__ P0, TransferMap tm,
                    double* __restrict__ rc, int64_t n, int out) {
    __shared__ double rs[M][B + 8];
    const int t = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * out + t;
    const bool active = e < n && t < out;
    double r[M];
    if (e < n) {
        const double* T = mat + (e >> 5) * (int64_t)(K * AMG1D_TILE) + (e & 31);
        double xl[M], xc[M], xr[M], y[M], bb[M];
        load_vec<M>(b + e * M, bb);
        load_vec<M>(x + (e - 1) * M, xl);
        load_vec<M>(x + e * M, xc);
        load_vec<M>(x + (e + 1) * M, xr);
        stream_Ax<M>(T, xl, xc, xr, y);
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

// ---- host-side dispatch ---------------------------------------------------------------------------------
#define FUSED_FOR_M(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)

inline bool fused_sweep(int m, int diag, const double* mat, const double* b, const double* xin,
                        double* xout, int64_t n, double alpha, int zero_guess, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    switch (m * 2 + (diag ? 1 : 0)) {
#define X(MM)                                                                                   \
    case MM * 2: f_sweep<MM, false><<<grid, 256, 0, st>>>(mat, b, xin, xout, n, alpha, zero_guess); return true; \
    case MM * 2 + 1: f_sweep<MM, true><<<grid, 256, 0, st>>>(mat, b, xin, xout, n, alpha, zero_guess); return true;
        FUSED_FOR_M(X)
#undef X
        default: return false;
    }
}

inline bool fused_resnorm(int m, int diag, const double* mat, const double* b, const double* x,
                          int64_t n, double* partial, int64_t partial_cap, int* nblocks,
                          cudaStream_t st) {
    const int64_t grid = (n + 255) / 256;
    if (grid > partial_cap) return false;
    *nblocks = (int)grid;
    switch (m * 2 + (diag ? 1 : 0)) {
#define X(MM)                                                                                                   \
    case MM * 2: f_resnorm<MM, 4 * MM * MM><<<(unsigned)grid, 256, 0, st>>>(mat, b, x, n, partial); return true;   \
    case MM * 2 + 1: f_resnorm<MM, 3 * MM * MM + MM><<<(unsigned)grid, 256, 0, st>>>(mat, b, x, n, partial); return true;
        FUSED_FOR_M(X)
#undef X
        default: return false;
    }
}

// (M, MC) pairs with a register-resident multi-sweep kernel
#define FUSED_PAIRS(X) X(1, 1) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(4, 2) X(4, 3) X(5, 3)

inline int fused_out_per_cta(int nsweep, int ratio) {
    const int out = ((FUSED_B - 2 * (nsweep + 1)) / ratio) * ratio;
    return out;
}

inline bool fused_down(int m, int mc, int diag, const TransferMap& tm, int nsweep, bool zero,
                       const double* mat, const double* b, const double* xin, double* xout,
                       const double* P0, double* rc, int64_t n, double alpha, const Slab& sl,
                       cudaStream_t st) {
    if (diag) return false;
    const int out = fused_out_per_cta(nsweep, tm.ratio);
    if (out < tm.ratio || out < FUSED_B / 2) return false;
    const unsigned grid = (unsigned)((n + out - 1) / out);
    switch (m * 16 + mc) {
#define X(MM, MCC)                                                                                     \
    case MM * 16 + MCC:                                                                                \
        f_down<MM, MCC, FUSED_B><<<grid, FUSED_B, 0, st>>>(mat, b, xin, xout, P0, tm, rc, n, alpha,   \
                                                            nsweep, zero ? 1 : 0, out, sl);           \
        return true;
        FUSED_PAIRS(X)
#undef X
        default: return false;
    }
}

inline bool fused_up(int m, int mc, int diag, const TransferMap& tm, int nsweep, const double* mat,
                     const double* b, const double* xin, double* xout, const double* P0,
                     const double* xcoarse, int64_t n, double alpha, double* partial,
                     int64_t partial_cap, int* nblocks, const Slab& sl, cudaStream_t st) {
    if (diag) return false;
    const int out = fused_out_per_cta(nsweep, tm.ratio);
    if (out < tm.ratio || out < FUSED_B / 2) return false;
    const int64_t grid = (n + out - 1) / out;
    if (partial && grid > partial_cap) return false;
    if (nblocks) *nblocks = (int)grid;
    switch (m * 16 + mc) {
#define X(MM, MCC)                                                                                      \
    case MM * 16 + MCC:                                                                                 \
        f_up<MM, MCC, FUSED_B><<<(unsigned)grid, FUSED_B, 0, st>>>(mat, b, xin, xout, P0, tm, xcoarse, \
                                                                    n, alpha, nsweep, out, partial, sl); \
        return true;
        FUSED_PAIRS(X)
#undef X
        default: return false;
    }
}

inline bool fused_residual_restrict(int m, int mc, int K, const TransferMap& tm, const double* mat,
                                    const double* b, const double* x, const double* P0, double* rc,
                                    int64_t n, cudaStream_t st) {
    const int out = (FUSED_B / tm.ratio) * tm.ratio;
    if (out < tm.ratio) return false;
    const unsigned grid = (unsigned)((n + out - 1) / out);
    switch (m * 16 + mc) {
#define X(MM, MCC)                                                                                   \
    case MM * 16 + MCC:                                                                              \
        f_residual_restrict<MM, MCC, FUSED_B><<<grid, FUSED_B, 0, st>>>(mat, K, b, x, P0, tm, rc, n, out); \
        return true;
        FUSED_PAIRS(X)
        X(9, 5) X(5, 2) X(9, 2)
#undef X
        default: return false;
    }
}
