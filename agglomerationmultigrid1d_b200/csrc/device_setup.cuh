// Device-side set-up of the DG-type level chain: what the reference does on the host in
//   src/mesh_heirarchy.jl:75-106, :160-176   G, D, C  <-  L' (G, D, C) L  (Galerkin products, sparse)
//                                            A = C - D (M \ G)            (block-diagonal mass matrix)
//   src/smoother.jl:154-164                  block-Jacobi blocks A[el, el] and their factorisation
// in element-block form on the GPU.  G, D, C are block tridiagonal in the element grouping, the
// transfers of the DG-type chain (dg_dg, aggdg_dg, aggdg_aggdg) are element local with a single parent,
// so every product is a small dense computation per (coarse) element:
//
//   Xc_di[K] = sum_{e in ch(K)} P[e]' ( X_di[e] P[e] + [e-1 in ch(K)] X_lo[e] P[e-1] + [e+1 in ch(K)] X_up[e] P[e+1] )
//   Xc_lo[K] = P[e0]' X_lo[e0] P[e0-1]        e0 = first child of K (e0 - 1 = last child of K - 1)
//   Xc_up[K] = P[e1]' X_up[e1] P[e1+1]        e1 = last child of K
//
//   A_di[e] = C_di[e] - ( D_lo[e] Mi[e-1] G_up[e-1] + D_di[e] Mi[e] G_di[e] + D_up[e] Mi[e+1] G_lo[e+1] )
//   A_lo[e] = C_lo[e] - ( D_lo[e] Mi[e-1] G_di[e-1] + D_di[e] Mi[e] G_lo[e] )
//   A_up[e] = C_up[e] - ( D_di[e] Mi[e] G_up[e]     + D_up[e] Mi[e+1] G_di[e+1] )
// with Mi = M^-1 per element.  The product D M^-1 G is block PENTA-diagonal in general; the reference's
// one-sided fluxes make its outer bands vanish (every level operator is block tridiagonal),
// which the kernel verifies instead of assuming (band_max reports the largest outer-band entry).
// All arrays here are plain element-block arrays (n blocks of m x m, column-major); k_repack then
// brings the result into the tile layout of the solver kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels_generic.cuh"
#include "layout.cuh"

// entry (i, j) of  P[ea]' X P[eb]  for X = X[e] (m_f x m_f), P blocks m_f x m_c column-major
__device__ __forceinline__ double ds_ptxp(const double* Pa, const double* X, const double* Pb, int mf,
                                          int i, int j) {
    double s = 0.0;
    for (int b = 0; b < mf; ++b) {
        double t = 0.0;
        for (int a = 0; a < mf; ++a) t = fma(Pa[i * mf + a], X[b * mf + a], t);   // (P' X)(i, b)
        s = fma(t, Pb[j * mf + b], s);
    }
    return s;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// one thread per (coarse element K, column j, row i); single-parent transfers only
__global__ void k_galerkin_tri(const double* __restrict__ lo, const double* __restrict__ di,
                               const double* __restrict__ up, const double* __restrict__ P, TransferMap tm,
                               int mf, int mc, double* __restrict__ lo_c, double* __restrict__ di_c,
                               double* __restrict__ up_c) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int mmc = mc * mc, mmf = mf * mf, bs = mf * mc;
    if (t >= tm.n_coarse * mmc) return;
    const int64_t K = t / mmc;
    const int q = (int)(t % mmc);
    const int j = q / mc, i = q % mc;
    const int64_t c0 = tm.first(K), c1 = tm.first(K + 1);
    double d = 0.0, l = 0.0, u = 0.0;
    for (int64_t e = c0; e < c1; ++e) {
        const double* Pe = P + tm.blk(e) * bs;
        d += ds_ptxp(Pe, di + e * mmf, Pe, mf, i, j);
        if (e > c0) d += ds_ptxp(Pe, lo + e * mmf, P + tm.blk(e - 1) * bs, mf, i, j);
        if (e < c1 - 1) d += ds_ptxp(Pe, up + e * mmf, P + tm.blk(e + 1) * bs, mf, i, j);
    }
    if (c1 > c0) {
        if (c0 > 0) l = ds_ptxp(P + tm.blk(c0) * bs, lo + c0 * mmf, P + tm.blk(c0 - 1) * bs, mf, i, j);
        if (c1 < tm.n_fine) u = ds_ptxp(P + tm.blk(c1 - 1) * bs, up + (c1 - 1) * mmf, P + tm.blk(c1) * bs, mf, i, j);
    }
    di_c[t] = d;
    lo_c[t] = l;
    up_c[t] = u;
}

// Galerkin product of the stiffness matrix itself, A_c = L' A L, for any element-local transfer: one
// parent (L[e, par(e)] = P0[e]) or two (additionally L[e, par(e) + 1] = P1[e]: cg_cg / dg_cg / aggdg_cg,
// where the nodes of a fine group interpolate from the coarse group of their element and from the
// vertex that opens the next one).  This is the CG loop of the reference's first constructor,
// mStiffness[i] = L' mStiffness[i-1] L (src/mesh_heirarchy.jl:52-60), in element-block form:
//
//   A_c[K, J] = sum over fine e with L[e, K] != 0, f in {e-1, e, e+1} with L[f, J] != 0 of
//               L[e, K]' A[e, f] L[f, J]
//
// One thread per (coarse element K, column j, row i) accumulates J = K-1, K, K+1; contributions to any
// other J (the product of two-parent transfers is block PENTA-diagonal in general; the reference's
// nodal interpolation makes the outer bands vanish) go to band_max[0] and are checked by the host.
// band_max[1] <- max |diagonal-block entry|.
__global__ void k_galerkin_general(const double* __restrict__ lo, const double* __restrict__ di,
                                   const double* __restrict__ up, const double* __restrict__ P0,
                                   const double* __restrict__ P1, TransferMap tm, int mf, int mc, int64_t nc,
                                   double* __restrict__ lo_c, double* __restrict__ di_c,
                                   double* __restrict__ up_c, double* __restrict__ band_max) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int mmc = mc * mc, mmf = mf * mf, bs = mf * mc;
    if (t >= nc * mmc) return;
    const int64_t K = t / mmc;
    const int q = (int)(t % mmc);
    const int j = q / mc, i = q % mc;
    const int64_t e0 = P1 ? tm.first(K - 1) : tm.first(K), e1 = tm.first(K + 1);
    double d = 0.0, l = 0.0, u = 0.0, band = 0.0;
    for (int64_t e = e0; e < e1; ++e) {
        const double* Pe = (tm.par(e) == K ? P0 : P1) + tm.blk(e) * bs;        // L[e, K]
        for (int s = -1; s <= 1; ++s) {
            const int64_t f = e + s;
            if (f < 0 || f >= tm.n_fine) continue;
            const double* X = (s < 0 ? lo : (s == 0 ? di : up)) + e * mmf;     // A[e, f]
            const int64_t pf = tm.par(f);
            for (int two = 0; two < (P1 ? 2 : 1); ++two) {
                const int64_t J = pf + two;                                    // L[f, J]
                const double v = ds_ptxp(Pe, X, (two ? P1 : P0) + tm.blk(f) * bs, mf, i, j);
                if (J == K) d += v;
                else if (J == K - 1) l += v;
                else if (J == K + 1 && J < nc) u += v;
                else band = fmax(band, fabs(v));
            }
        }
    }
    di_c[t] = d;
    lo_c[t] = K > 0 ? l : 0.0;
    up_c[t] = K < nc - 1 ? u : 0.0;
    if (K == 0 && l != 0.0) band = fmax(band, fabs(l));
    if (band > 0.0) atomic_max_nonneg(band_max, band);
    atomic_max_nonneg(band_max + 1, fabs(d));
}

// padding slots of a regrouped (CG) level, perm[slot] < 0: identity on the diagonal, as the host's
// block extraction does (blocks.csc_to_blocks)
__global__ void k_pad_identity(double* __restrict__ di, const int64_t* __restrict__ perm, int64_t n, int m) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n * m) return;
    if (perm[s] < 0) {
        const int64_t e = s / m;
        const int i = (int)(s % m);
        di[e * m * m + i * m + i] = 1.0;
    }
}

// point-Jacobi smoother (src/smoother.jl:92-98): dinv[e*m + i] = 1 / A_di[e](i, i)
__global__ void k_diag_reciprocal(const double* __restrict__ di, int64_t n, int m, double* __restrict__ dinv,
                                  int* __restrict__ flag) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n * m) return;
    const int64_t e = s / m;
    const int i = (int)(s % m);
    const double a = di[e * m * m + i * m + i];
    if (a == 0.0) { flag[0] = 1; dinv[s] = 0.0; return; }
    dinv[s] = 1.0 / a;
}

// entry (i, j) of  D Mi G  (all m x m column-major)
__device__ __forceinline__ double ds_dmg(const double* D, const double* Mi, const double* G, int m, int i, int j) {
    double s = 0.0;
    for (int b = 0; b < m; ++b) {
        double t = 0.0;
        for (int a = 0; a < m; ++a) t = fma(D[a * m + i], Mi[b * m + a], t);       // (D Mi)(i, b)
        s = fma(t, G[j * m + b], s);
    }
    return s;
}


// A = C - D M^-1 G, tridiagonal part; band_max[0] <- max |outer-band entry|.  One thread per (e, j, i).
// Mi: n blocks, or one block for every element when mi_const != 0.
__global__ void k_flux_operator(const double* __restrict__ Glo, const double* __restrict__ Gdi,
                                const double* __restrict__ Gup, const double* __restrict__ Dlo,
                                const double* __restrict__ Ddi, const double* __restrict__ Dup,
                                const double* __restrict__ Clo, const double* __restrict__ Cdi,
                                const double* __restrict__ Cup, const double* __restrict__ Mi, int mi_const,
                                int64_t n, int m, double* __restrict__ Alo, double* __restrict__ Adi,
                                double* __restrict__ Aup, double* __restrict__ band_max) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int mm = m * m;
    if (t >= n * mm) return;
    const int64_t e = t / mm;
    const int q = (int)(t % mm);
    const int j = q / m, i = q % m;
    auto mi = [&](int64_t k) { return Mi + (mi_const ? 0 : k * mm); };
    const bool hl = e > 0, hr = e < n - 1;
    double d = ds_dmg(Ddi + e * mm, mi(e), Gdi + e * mm, m, i, j);
    double l = ds_dmg(Ddi + e * mm, mi(e), Glo + e * mm, m, i, j);
    double u = ds_dmg(Ddi + e * mm, mi(e), Gup + e * mm, m, i, j);
    double band = 0.0;
    if (hl) {
        d += ds_dmg(Dlo + e * mm, mi(e - 1), Gup + (e - 1) * mm, m, i, j);
        l += ds_dmg(Dlo + e * mm, mi(e - 1), Gdi + (e - 1) * mm, m, i, j);
        band = fmax(band, fabs(ds_dmg(Dlo + e * mm, mi(e - 1), Glo + (e - 1) * mm, m, i, j)));
    }
    if (hr) {
        d += ds_dmg(Dup + e * mm, mi(e + 1), Glo + (e + 1) * mm, m, i, j);
        u += ds_dmg(Dup + e * mm, mi(e + 1), Gdi + (e + 1) * mm, m, i, j);
        band = fmax(band, fabs(ds_dmg(Dup + e * mm, mi(e + 1), Gup + (e + 1) * mm, m, i, j)));
    }
    Adi[t] = Cdi[t] - d;
    Alo[t] = hl ? Clo[t] - l : 0.0;
    Aup[t] = hr ? Cup[t] - u : 0.0;
    if (band > 0.0) atomic_max_nonneg(band_max, band);
    atomic_max_nonneg(band_max + 1, fabs(Adi[t]));
}

// Union sparsity masks of the off-diagonal blocks: masks[0..m) columns of lo, [m..2m) rows of lo,
// [2m..3m) columns of up, [3m..4m) rows of up.
__global__ void k_structure_masks(const double* __restrict__ lo, const double* __restrict__ up, int64_t n,
                                  int m, int* __restrict__ masks) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int mm = m * m;
    if (t >= n * mm) return;
    const int q = (int)(t % mm);
    const int j = q / m, i = q % m;
    if (lo[t] != 0.0) { masks[j] = 1; masks[m + i] = 1; }
    if (up[t] != 0.0) { masks[2 * m + j] = 1; masks[3 * m + i] = 1; }
}

// Dinv rows of a level's tiles -> element-block array (n blocks of m*m, or n * m for a diagonal smoother)
__global__ void k_extract_dinv(const double* __restrict__ mat, MatDesc d, int64_t n, double* __restrict__ dinv) {
    const int dsz = d.diag ? d.m : d.m * d.m;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * dsz) return;
    const int64_t e = t / dsz;
    const int q = (int)(t % dsz);
    dinv[t] = mat[(e >> 5) * (int64_t)d.K * AMG1D_TILE + (int64_t)(d.o_dv + q) * AMG1D_TILE + (e & 31)];
}

// ---- device-computed smoother inverses ("recompute_dinv") -------------------------------------------------
// The block-Jacobi inverse of an element is a function of its diagonal block A_di, which the fused legs hold in
// registers anyway: they can invert it in-kernel instead of streaming 16 of the 40 stored doubles of a 4 x 4
// DG element from HBM (kernels_fused.cuh: reg_invert).  For the fused legs and every kernel that READS the
// stored inverse (generic tier, streaming kernels, single-CTA tail, pattern tables) to stay bit-identical, the
// stored inverse must be THE SAME Gauss-Jordan result.  This kernel recomputes it per element with the
// arithmetic of reg_invert (kernels_fused.cuh: partial pivoting by a compare-and-swap chain, same operation order).
// (k_dinv_recompute below = the kernel template k_dinv_recompute_t behind launch_dinv_recompute.)
//   mode 0: dev[0] = max over the elements of  max|Dinv_stored - inv(A_di)| / max|inv(A_di)|  (as an ordered
//           int64 bit pattern, atomicMax), flag[0] |= 1 for a singular block, |= 2 if any element swaps rows,
//           |= 4 if some element that swaps rows NEEDS to (below) - nothing is written;
//   mode 1: overwrite the stored Dinv rows with the recomputed (pivoted) inverse;
//   mode 2: modes 0 and 1 in one pass over the level - an element is overwritten iff ITS upload agrees to 1e-8 (if
//           some element of the level fails, the level is not marked for recomputation and keeps a stored inverse
//           that is valid element by element);
//   mode 3: overwrite the stored Dinv rows with the UNPIVOTED Gauss-Jordan inverse (reg_invert<M, false>).
// Pivoting is a property of the node numbering more than of the numbers: the DG blocks of the reference (end nodes
// first, src/dg_mesh.jl:41-46) swap rows in every element although they are symmetric positive definite, for which
// elimination without pivoting is just as stable - and the swap chain is a third of the instructions of a fused leg
// that inverts in registers (192 selects per 4 x 4 element).  So for every element that swaps, modes 0 / 2 also run
// the elimination WITHOUT pivoting and compare: if the two inverses agree to 1e-12 (relative to the largest entry) in
// every such element, the level adopts the unpivoted inverse (host: a second pass in mode 3) and its legs skip the
// chain; one element that disagrees, or hits a zero pivot, keeps the whole level on the pivoted path (flag 4).
// Addressing: element e (e_first <= e < e_end, may be negative: left ghosts) lives at
// base + (e >> 5) * K * tile_stride + (e & 31) with row stride tile_stride (element tiles: tile_stride = 32);
// a pattern table tab[set][k] is addressed with tile_stride = 1 and "elements" = its rows: base + e * K.
#define AMG1D_DVREC_MAXM 9
#define AMG1D_NOPIVOT_TOL 1e-12
// The kernel is instantiated for the block sizes 1 .. 5 (MT = m: every loop unrolls, the block lives in registers) and
// once for any size up to AMG1D_DVREC_MAXM (MT = 0: run-time loops over per-thread arrays, i.e. local memory - 160 ms
// of T's set-up before the instantiations existed); the statements, and with them the bits, are the same.
// in-place Gauss-Jordan inverse without pivoting (column-major m x m), statement for statement reg_invert<M, false>
template <int MT>
__device__ __forceinline__ bool gj_invert_nopivot(double* A, int m_rt) {
    const int m = MT > 0 ? MT : m_rt;
#pragma unroll
    for (int c = 0; c < m; ++c) {
        if (A[c * m + c] == 0.0) return false;
        const double dd = 1.0 / A[c * m + c];
        A[c * m + c] = 1.0;
#pragma unroll
        for (int q = 0; q < m; ++q) A[q * m + c] *= dd;
#pragma unroll
        for (int r = 0; r < m; ++r) {
            if (r == c) continue;
            const double f = A[c * m + r];
            A[c * m + r] = 0.0;
#pragma unroll
            for (int q = 0; q < m; ++q) A[q * m + r] = fma(-f, A[q * m + c], A[q * m + r]);
        }
    }
    return true;
}

template <int MT>
__global__ void k_dinv_recompute_t(double* __restrict__ base, MatDesc d, int64_t e_first, int64_t e_end, int tile_stride,
                                   int mode, unsigned long long* __restrict__ dev, int* __restrict__ flag) {
    constexpr int MM = MT > 0 ? MT : AMG1D_DVREC_MAXM;
    const int64_t e = e_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_end) return;
    const int m = MT > 0 ? MT : d.m;
    double* T = tile_stride == 1 ? base + e * (int64_t)d.K
                                 : base + (e >> 5) * (int64_t)d.K * AMG1D_TILE + (e & 31);    // e >> 5 floors
    const int rs = tile_stride == 1 ? 1 : AMG1D_TILE;
    double A[MM * MM];
    bool sw[MM][MM];
    bool allzero = true;
#pragma unroll
    for (int k = 0; k < m * m; ++k) { A[k] = T[(int64_t)(d.o_di + k) * rs]; allzero = allzero && A[k] == 0.0; }
    if (allzero) return;                          // zero-filled slots outside the level (slab padding)
    if (mode == 3) {
        if (!gj_invert_nopivot<MT>(A, m)) { atomicOr(flag, 1); return; }
#pragma unroll
        for (int k = 0; k < m * m; ++k) T[(int64_t)(d.o_dv + k) * rs] = A[k];
        return;
    }
    bool swapped = false;
#pragma unroll
    for (int c = 0; c < m; ++c) {
#pragma unroll
        for (int r = c + 1; r < m; ++r) {         // compare-and-swap chain: the largest |a(r, c)|, r >= c, ends on the diagonal
            const bool s = fabs(A[c * m + r]) > fabs(A[c * m + c]);
            sw[c][r] = s;
            swapped = swapped || s;
#pragma unroll
            for (int q = 0; q < m; ++q) {
                const double x = A[q * m + c], y = A[q * m + r];
                A[q * m + c] = s ? y : x;
                A[q * m + r] = s ? x : y;
            }
        }
        if (A[c * m + c] == 0.0) { atomicOr(flag, 1); return; }
        const double dd = 1.0 / A[c * m + c];
        A[c * m + c] = 1.0;
#pragma unroll
        for (int q = 0; q < m; ++q) A[q * m + c] *= dd;
#pragma unroll
        for (int r = 0; r < m; ++r) {
            if (r == c) continue;
            const double f = A[c * m + r];
            A[c * m + r] = 0.0;
#pragma unroll
            for (int q = 0; q < m; ++q) A[q * m + r] = fma(-f, A[q * m + c], A[q * m + r]);
        }
    }
#pragma unroll
    for (int c = m - 1; c >= 0; --c) {            // undo the row swaps as column swaps, in reverse order
#pragma unroll
        for (int r2 = m - 1; r2 > c; --r2) {
            const bool s = sw[c][r2];
#pragma unroll
            for (int r = 0; r < m; ++r) {
                const double x = A[c * m + r], y = A[r2 * m + r];
                A[c * m + r] = s ? y : x;
                A[r2 * m + r] = s ? x : y;
            }
        }
    }
    if (mode != 1) {
        double mx = 0.0, df = 0.0;
#pragma unroll
        for (int k = 0; k < m * m; ++k) {
            mx = fmax(mx, fabs(A[k]));
            df = fmax(df, fabs(A[k] - T[(int64_t)(d.o_dv + k) * rs]));
        }
        const double rel = mx > 0.0 ? df / mx : 0.0;
        atomicMax(dev, (unsigned long long)__double_as_longlong(rel >= 0.0 ? rel : INFINITY));   // NaN -> inf
        if (swapped) {                            // does this element need its pivots?
            atomicOr(flag, 2);
            double U[MM * MM];
#pragma unroll
            for (int k = 0; k < m * m; ++k) U[k] = T[(int64_t)(d.o_di + k) * rs];
            bool same = gj_invert_nopivot<MT>(U, m);
            double du = 0.0;
#pragma unroll
            for (int k = 0; k < m * m; ++k) du = fmax(du, fabs(U[k] - A[k]));
            same = same && (du <= AMG1D_NOPIVOT_TOL * mx);                                    // false for NaN
            if (!same) atomicOr(flag, 4);
        }
        if (mode == 2 && !(rel <= 1e-8)) return;      // this element's upload is not inv(A_di): leave it alone
    }
    if (mode != 0) {
#pragma unroll
        for (int k = 0; k < m * m; ++k) T[(int64_t)(d.o_dv + k) * rs] = A[k];
    }
}

inline void launch_dinv_recompute(unsigned grid, cudaStream_t st, double* base, const MatDesc& d, int64_t e_first,
                                  int64_t e_end, int tile_stride, int mode, unsigned long long* dev, int* flag) {
    switch (d.m) {
        case 1: k_dinv_recompute_t<1><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
        case 2: k_dinv_recompute_t<2><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
        case 3: k_dinv_recompute_t<3><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
        case 4: k_dinv_recompute_t<4><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
        case 5: k_dinv_recompute_t<5><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
        default: k_dinv_recompute_t<0><<<grid, 128, 0, st>>>(base, d, e_first, e_end, tile_stride, mode, dev, flag); break;
    }
}

// ---- device-side right-hand side ----------------------------------------------------------------
// The volume part of dg_flux_rhs (src/dg_mesh.jl:342-365: f[el.mNodesInd] += J sum_q w_q phi_i(xi_q) func(x_q)) and of
// cg_stiffness_and_rhs / cg_rhs (src/cg_mesh.jl:150-160, :205-215), with func given as a sum of terms
//     coef * x^pow * g(w x + phi),   g = 1 | cos | sin | exp            (terms[5 t ..]: kind, coef, pow, w, phi)
// - the manufactured right-hand sides of the reference's scripts (cos(x), exp(-x), 1) and of the BASELINE configs.
// One thread per element: x_q = xc + (h / 2) xi_q on the element [xl, xr], xl = xin + (e / n)(xout - xin) exactly as
// tests/mesh_generator.jl:20-32 computes the vertices (or xl, xr from an uploaded vertex array), then
// fe[i] = (h / 2) sum_q func(x_q) W[q][i], W[q][i] = w_q phi_i(xi_q).
//   kind 0 (DG-type level):  b[e m + i] = fe[i]
//   kind 1 (CG level in group form, group k = [vertex k, interior nodes of element k]): vertex k += fe[0] of element
//          k and fe[1] of element k - 1; interior slot j = fe[j + 1].  Two contributions per vertex onto a zeroed b:
//          the sum does not depend on their order.
// e runs over the rank's slab (global element e_off + e); b points at local element / group 0.
#define AMG1D_RHS_MAXQ 16
#define AMG1D_RHS_MAXM 10
#define AMG1D_RHS_MAXT 8
struct RhsSpec {
    int kind, nq, m, n_terms;                 // m = local basis functions per element (p + 1)
    double xi[AMG1D_RHS_MAXQ];
    double W[AMG1D_RHS_MAXQ * AMG1D_RHS_MAXM];  // [q][i]
    double terms[5 * AMG1D_RHS_MAXT];
    double xin, xout;
    int64_t n_glob;                           // elements of the whole mesh
};

__device__ __forceinline__ double rhs_func(const RhsSpec& sp, double x) {
    double s = 0.0;
    for (int t = 0; t < sp.n_terms; ++t) {
        const double* T = sp.terms + 5 * t;
        const int kind = (int)T[0];
        const double arg = fma(T[3], x, T[4]);
        double g = kind == 1 ? cos(arg) : kind == 2 ? sin(arg) : kind == 3 ? exp(arg) : 1.0;
        const int pw = (int)T[2];
        for (int k = 0; k < pw; ++k) g *= x;
        s = fma(T[1], g, s);
    }
    return s;
}

// only_next = 1 (kind 1): add nothing but the element's fe[1] to the NEXT group's vertex - for the element left of a
// slab's first ghost group, whose own group lies outside the slab.
__global__ void k_assemble_rhs(const __grid_constant__ RhsSpec sp, const double* __restrict__ vertices, int64_t e_off,
                               int64_t e_begin, int64_t e_end, int64_t slots_end, double* __restrict__ b, int only_next) {
    const int64_t e = e_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // local element (may be a ghost)
    if (e >= e_end) return;
    const int64_t eg = e + e_off;
    if (eg < 0 || eg >= sp.n_glob) return;
    double xl, xr;
    if (vertices) { xl = vertices[eg]; xr = vertices[eg + 1]; }
    else {
        const double L = __dsub_rn(sp.xout, sp.xin), nn = (double)sp.n_glob;
        xl = __dadd_rn(sp.xin, __dmul_rn(__ddiv_rn((double)eg, nn), L));
        xr = __dadd_rn(sp.xin, __dmul_rn(__ddiv_rn((double)(eg + 1), nn), L));
    }
    const double hh = __dsub_rn(xr, xl), xc = __dmul_rn(__dadd_rn(xl, xr), 0.5), hj = __dmul_rn(hh, 0.5);
    double fe[AMG1D_RHS_MAXM];
    for (int i = 0; i < sp.m; ++i) fe[i] = 0.0;
    for (int q = 0; q < sp.nq; ++q) {
        const double fx = rhs_func(sp, __dadd_rn(xc, __dmul_rn(hj, sp.xi[q])));
        for (int i = 0; i < sp.m; ++i) fe[i] = fma(fx, sp.W[q * AMG1D_RHS_MAXM + i], fe[i]);
    }
    if (sp.kind == 0) {
        for (int i = 0; i < sp.m; ++i) b[e * sp.m + i] = __dmul_rn(hj, fe[i]);
    } else {
        const int p = sp.m - 1;                              // group size
        if (!only_next) {
            atomicAdd(&b[e * p], __dmul_rn(hj, fe[0]));
            for (int j = 1; j < p; ++j) b[e * p + j] = __dmul_rn(hj, fe[j + 1]);
        }
        if ((e + 1) * p < slots_end) atomicAdd(&b[(e + 1) * p], __dmul_rn(hj, fe[1]));
    }
}

// ops[k] = 0: b[slot] += val;  1: b[slot] = val   (boundary terms; applied in order by ONE thread - a handful of entries)
__global__ void k_apply_fixes(double* __restrict__ b, int64_t slot_off, int64_t slot_begin, int64_t slot_end, int n,
                              const int64_t* __restrict__ slots, const double* __restrict__ vals,
                              const int* __restrict__ ops) {
    if (blockIdx.x || threadIdx.x) return;
    for (int k = 0; k < n; ++k) {
        const int64_t s = slots[k] - slot_off;               // local slot
        if (s < slot_begin || s >= slot_end) continue;
        b[s] = ops[k] ? vals[k] : b[s] + vals[k];
    }
}
