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
