"""Oracle (test infrastructure): block-diagonal matrix container and its LU.

Follows src/block_diagonal.jl:11-21 (types), :27-76 (constructors), :166-176 (``mul!``),
:195-274 (block-diagonal x sparse), :299-309 (``ldiv!``), :314-393 (LU \\ sparse).
Note the reference's ``mul!`` / ``ldiv!`` ACCUMULATE into their output (``+=``, :172, :305); the
``*`` and ``\\`` operators pass a zero output so the net result is the plain product / solve.
Block indices (mBlockInds, one column per block) are 0-based here.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


class BlockDiagonal:
    def __init__(self, mBlocks, mBlockSize=None, mBlockInds=None):
        self.mBlocks = [np.array(b, dtype=np.float64) for b in mBlocks]
        if mBlockSize is None:
            # src/block_diagonal.jl:27-41
            mBlockSize = self.mBlocks[0].shape[0]
            mBlockInds = np.zeros((mBlockSize, len(self.mBlocks)), dtype=np.int64)
            for i, blk in enumerate(self.mBlocks):
                mBlockInds[:, i] = np.arange(i * mBlockSize, (i + 1) * mBlockSize)
                if blk.shape != (mBlockSize, mBlockSize):
                    raise ValueError("All blocks must be of the same size.")
        self.mBlockSize = mBlockSize
        self.mBlockInds = np.asarray(mBlockInds, dtype=np.int64)

    @property
    def shape(self):
        n = len(self.mBlocks) * self.mBlockSize
        return (n, n)

    def toarray(self):
        B = np.zeros(self.shape)
        for i, blk in enumerate(self.mBlocks):
            idx = self.mBlockInds[:, i]
            B[np.ix_(idx, idx)] = blk
        return B

    def tosparse(self):
        rows, cols, vals = [], [], []
        for i, blk in enumerate(self.mBlocks):
            idx = self.mBlockInds[:, i]
            for b in range(self.mBlockSize):
                for a in range(self.mBlockSize):
                    rows.append(idx[a]); cols.append(idx[b]); vals.append(blk[a, b])
        return sp.csc_matrix((vals, (rows, cols)), shape=self.shape)

    def mul_into(self, C, B):
        """``mul!(C, A, B)``: C[inds, :] += block * B[inds, :] (src/block_diagonal.jl:166-176)."""
        for i, blk in enumerate(self.mBlocks):
            idx = self.mBlockInds[:, i]
            C[idx] += blk @ B[idx]
        return C

    def __matmul__(self, B):
        if sp.issparse(B):
            return bd_sp_matmul(self, B)
        B = np.asarray(B, dtype=np.float64)
        return self.mul_into(np.zeros(B.shape), B)

    def lu(self):
        return BlockDiagonalLU(self)


class BlockDiagonalLU:
    """src/block_diagonal.jl:17-21, :47-58: one partial-pivoting LU (LAPACK getrf) per block."""

    def __init__(self, A):
        self.mBlockSize = A.mBlockSize
        self.mBlockInds = A.mBlockInds
        self.mBlocksLU = [sla.lu_factor(blk) for blk in A.mBlocks]

    @property
    def shape(self):
        n = len(self.mBlocksLU) * self.mBlockSize
        return (n, n)

    def ldiv_into(self, C, B):
        """``ldiv!(C, A, B)``: C[inds, :] += LU_i \\ B[inds, :] (src/block_diagonal.jl:299-309)."""
        for i, lu in enumerate(self.mBlocksLU):
            idx = self.mBlockInds[:, i]
            C[idx] += sla.lu_solve(lu, B[idx])
        return C

    def solve(self, B):
        """``A \\ B`` for dense vectors / matrices and for sparse matrices."""
        if sp.issparse(B):
            return bd_sp_solve(self, B)
        B = np.asarray(B, dtype=np.float64)
        return self.ldiv_into(np.zeros(B.shape), B)


def _blockwise_sparse(apply_block, nblocks, inds, B):
    """Shared walk for block-diagonal (x or \\) sparse: each block only touches the columns in
    which its rows of B are non-zero (the reference walks column by column,
    src/block_diagonal.jl:195-274, :314-393; same arithmetic per (block, column))."""
    Bc = sp.csr_matrix(B)
    rows, cols, vals = [], [], []
    for i in range(nblocks):
        idx = inds[:, i]
        sub = Bc[idx, :].tocsc()
        nzcols = np.flatnonzero(np.diff(sub.indptr))
        if len(nzcols) == 0:
            continue
        dense = sub[:, nzcols].toarray()
        out = apply_block(i, dense)
        for b, c in enumerate(nzcols):
            for a in range(len(idx)):
                rows.append(idx[a]); cols.append(c); vals.append(out[a, b])
    return sp.csc_matrix((vals, (rows, cols)), shape=B.shape)


def bd_sp_matmul(A, B):
    return _blockwise_sparse(lambda i, d: A.mBlocks[i] @ d, len(A.mBlocks), A.mBlockInds, B)


def bd_sp_solve(A, B):
    return _blockwise_sparse(lambda i, d: sla.lu_solve(A.mBlocksLU[i], d), len(A.mBlocksLU),
                             A.mBlockInds, B)
