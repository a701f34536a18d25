// Device data layout shared by every kernel of libamg1d.
//
// Operator storage ("element tiles").  The host uploads each level in the element-block (BSR-like)
// form the C ABI documents; on the device the same numbers are kept interleaved in tiles of
// TILE = 32 consecutive elements:
//
//      mat[tile][k][lane]      tile = e / 32, lane = e % 32, k in [0, K)
//      k = 0*m*m + j*m + i  ->  A_lo[e](i,j)      (couples to element e-1)
//      k = 1*m*m + j*m + i  ->  A_di[e](i,j)
//      k = 2*m*m + j*m + i  ->  A_up[e](i,j)      (couples to element e+1)
//      k = 3*m*m + j*m + i  ->  Dinv[e](i,j)      (or 3*m*m + i for a diagonal smoother)
//
// so a warp that owns one tile streams one contiguous K*256-byte range and every warp-level load is
// a fully used, 256-byte coalesced request - for ANY block size m, with no shared-memory transpose
// and no bank conflicts.  The byte count is identical to the element-block layout.
//
// Vectors (x, b, r) stay element-major, v[e*m + i], exactly the reference's DG numbering
// (src/dg_mesh.jl:41-46), with one ghost element on each side of the slab: v[-m .. -1] and
// v[n*m .. n*m + m - 1] exist and hold zeros on a single GPU (A_lo[0] = A_up[n-1] = 0 there) or the
// neighbour rank's edge element in the sharded case.
#pragma once
#include <cstdint>

#define AMG1D_TILE 32

__host__ __device__ inline int amg1d_K(int m, int diag) { return 3 * m * m + (diag ? m : m * m); }

__host__ __device__ inline int64_t amg1d_tiles(int64_t n) { return (n + AMG1D_TILE - 1) / AMG1D_TILE; }
