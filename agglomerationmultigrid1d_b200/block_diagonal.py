"""Host set-up: block-diagonal operator type (mass matrices) and its LU.

Mirrors src/block_diagonal.jl:11-21 (types), :27-76 (constructors), :166-193 (products), :299-312
(solves), :195-274 / :314-393 (block-diagonal x sparse, LU \\ sparse).  Blocks are held as one
(n, m, m) array and applied as batched GEMV / batched LAPACK solves.  ``mul`` / ``ldiv`` keep the
reference's accumulate-into-C semantics (``C[inds,:] += ...``, :172, :305); ``@`` and ``solve`` start
from zeros.  mBlockInds (m, n) holds 0-based DOF ids, one column per block.
"""
import numpy as np
import scipy.sparse as sp


def _block_sparse(blocks, inds, shape):
    n, m, _ = blocks.shape
    rows = np.repeat(inds.T[:, :, None], m, axis=2)       # (n, i, j) -> row id
    cols = np.repeat(inds.T[:, None, :], m, axis=1)       # (n, i, j) -> col id
    return sp.csc_matrix((blocks.ravel(), (rows.ravel(), cols.ravel())), shape=shape)


class BlockDiagonal:
    __array_ufunc__ = None          # ndarray @ BlockDiagonal and ndarray * BlockDiagonal defer to __rmatmul__ / __rmul__

    def __init__(self, mBlocks, mBlockSize=None, mBlockInds=None):
        self.mBlocks = np.ascontiguousarray(mBlocks, dtype=np.float64)
        if self.mBlocks.ndim != 3 or self.mBlocks.shape[1] != self.mBlocks.shape[2]:
            raise ValueError("All blocks must be of the same size.")
        n, m, _ = self.mBlocks.shape
        self.mBlockSize = m if mBlockSize is None else int(mBlockSize)
        if mBlockInds is None:
            mBlockInds = np.arange(n * m, dtype=np.int64).reshape(n, m).T
        self.mBlockInds = np.ascontiguousarray(mBlockInds, dtype=np.int64)

    @property
    def shape(self):
        N = self.mBlocks.shape[0] * self.mBlockSize
        return (N, N)

    def toarray(self):
        return self.tosparse().toarray()

    def tosparse(self):
        return _block_sparse(self.mBlocks, self.mBlockInds, self.shape)

    def mul(self, Cout, B):
        """``mul!(C, A, B)``: C[inds, :] += block * B[inds, :]."""
        if B.shape[0] != self.shape[1] or Cout.shape != B.shape:
            raise ValueError("DimensionMismatch")
        idx = self.mBlockInds.T                                   # (n, m)
        Bb = B[idx]                                               # (n, m[, k])
        if B.ndim == 1:
            Cout[idx] += np.einsum("nij,nj->ni", self.mBlocks, Bb)
        else:
            Cout[idx] += np.einsum("nij,njk->nik", self.mBlocks, Bb)
        return Cout

    def __matmul__(self, B):
        if sp.issparse(B):
            return (self.tosparse() @ B).tocsc()
        B = np.asarray(B, dtype=np.float64)
        return self.mul(np.zeros(B.shape), B)

    def lu(self):
        return BlockDiagonalLU(self)

    # ---- the rest of the reference's surface (src/block_diagonal.jl:85-161, :181-193) -------------------
    def size(self, i=None):
        """``size(A)`` / ``size(A, i)`` (1-based i as in the reference: 1, 2 -> N; >= 3 -> 1)."""
        if i is None:
            return self.shape
        if i <= 0:
            raise ValueError("arraysize: dimension out of range")
        return 1 if i >= 3 else self.shape[0]

    def similar(self):
        """Same block structure, uninitialised blocks (src/block_diagonal.jl:95-105)."""
        return BlockDiagonal(np.empty_like(self.mBlocks), self.mBlockSize, self.mBlockInds)

    def scale(self, b, out=None):
        """``mul!(C, A, b)`` / ``mul!(C, b, A)``: C.blocks = b * A.blocks (:137-161)."""
        out = self.similar() if out is None else out
        if out.shape != self.shape:
            raise ValueError("DimensionMismatch")
        if out.mBlocks.shape[0] != self.mBlocks.shape[0]:
            raise ValueError("C and A must have the same number of blocks.")
        out.mBlockInds[:, :] = self.mBlockInds
        np.multiply(self.mBlocks, float(b), out=out.mBlocks)
        return out

    def __mul__(self, b):
        if np.isscalar(b):
            return self.scale(b)
        return self.__matmul__(b)

    def __rmul__(self, b):
        if np.isscalar(b):
            return self.scale(b)
        return NotImplemented

    def rmul(self, Cout, Adense):
        """``mul!(C, A', B)`` with B = self: C[:, inds] += A[:, inds] * block (:181-191)."""
        Adense = np.asarray(Adense, dtype=np.float64)
        if Adense.shape[1] != self.shape[0] or Cout.shape != (Adense.shape[0], self.shape[1]):
            raise ValueError("DimensionMismatch")
        idx = self.mBlockInds.T                                   # (n, m)
        Cout[:, idx] += np.einsum("rni,nij->rnj", Adense[:, idx], self.mBlocks)
        return Cout

    def __rmatmul__(self, Adense):
        Adense = np.asarray(Adense, dtype=np.float64)
        return self.rmul(np.zeros((Adense.shape[0], self.shape[1])), Adense)


class BlockDiagonalLU:
    """One partial-pivoting LU per block (the reference keeps ``la.lu`` factors; here the batched
    LAPACK ``gesv`` / ``getri`` of numpy does the same factorisation per block)."""

    def __init__(self, A):
        self.mBlockSize = A.mBlockSize
        self.mBlockInds = A.mBlockInds
        self._blocks = A.mBlocks
        self._inv = None

    @property
    def shape(self):
        N = self._blocks.shape[0] * self.mBlockSize
        return (N, N)

    @property
    def inverse_blocks(self):
        if self._inv is None:
            self._inv = np.linalg.inv(self._blocks)
        return self._inv

    def ldiv(self, Cout, B):
        """``ldiv!(C, A, B)``: C[inds, :] += LU_i \\ B[inds, :]."""
        if B.shape[0] != self.shape[1] or Cout.shape != B.shape:
            raise ValueError("DimensionMismatch")
        idx = self.mBlockInds.T
        Bb = B[idx]
        if B.ndim == 1:
            Cout[idx] += np.linalg.solve(self._blocks, Bb[:, :, None])[:, :, 0]
        else:
            Cout[idx] += np.linalg.solve(self._blocks, Bb)
        return Cout

    def solve(self, B):
        """``A \\ B`` for dense vectors / matrices and sparse matrices."""
        if sp.issparse(B):
            return (_block_sparse(self.inverse_blocks, self.mBlockInds, self.shape) @ B).tocsc()
        B = np.asarray(B, dtype=np.float64)
        return self.ldiv(np.zeros(B.shape), B)


def lu(A):
    return A.lu()


# ---- block-diagonal x sparse, column by column (src/block_diagonal.jl:195-274, :314-393) --------------------
def _touched_blocks(A, B, col):
    """Blocks of A hit by the stored entries of column `col` of the sparse B, and those entries scattered
    into one dense (n_touched, m) array - the reference's tempVec per block, for all blocks at once."""
    B = sp.csc_matrix(B)
    if B.shape[0] != A.shape[1]:
        raise ValueError("DimensionMismatch")
    lo, hi = B.indptr[col], B.indptr[col + 1]
    rows, vals = B.indices[lo:hi], B.data[lo:hi]
    m = A.mBlockSize
    blocks, inv = np.unique(rows // m, return_inverse=True)
    first = A.mBlockInds[0, blocks]                               # minBlockRow of every touched block
    rhs = np.zeros((len(blocks), m))
    rhs[inv, rows - first[inv]] = vals
    return blocks, rhs


def bd_sp_colmul(A, B, col):
    """(rowValCol, nzValCol) of column `col` (0-based) of A * B: every row of every touched block, stored
    zeros included, as the reference pushes them."""
    blocks, rhs = _touched_blocks(A, B, col)
    out = np.einsum("nij,nj->ni", A.mBlocks[blocks], rhs)
    return A.mBlockInds[:, blocks].T.ravel(), out.ravel()


def bd_sp_colsolve(A, B, col):
    """(rowValCol, nzValCol) of column `col` (0-based) of A \\ B for a BlockDiagonalLU A."""
    blocks, rhs = _touched_blocks(A, B, col)
    out = np.linalg.solve(A._blocks[blocks], rhs[:, :, None])[:, :, 0] if len(blocks) else rhs
    return A.mBlockInds[:, blocks].T.ravel(), out.ravel()


def _assemble_columns(A, B, colfun):
    B = sp.csc_matrix(B)
    ncol = B.shape[1]
    indptr = np.zeros(ncol + 1, dtype=np.int64)
    rows, vals = [], []
    for col in range(ncol):
        r, v = colfun(A, B, col)
        rows.append(r); vals.append(v)
        indptr[col + 1] = indptr[col] + len(r)
    return sp.csc_matrix((np.concatenate(vals) if vals else np.zeros(0),
                          np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64), indptr),
                         shape=(A.shape[0], ncol))


def bd_sp_matmul(A, B):
    """A * B for sparse B, assembled column by column; keeps the explicit zeros of touched blocks."""
    return _assemble_columns(A, B, bd_sp_colmul)


def bd_sp_solve(A, B):
    """A \\ B for a BlockDiagonalLU A and sparse B, column by column."""
    return _assemble_columns(A, B, bd_sp_colsolve)
