"""Oracle (test infrastructure): smoothers.

Follows src/smoother.jl:1-18 (additive Schwarz), :24-46 (hybrid Schwarz), :52-58 (point Jacobi),
:64-81 (block Jacobi), :88-139 (``cg_smoother``), :142-168 (``dg_smoother``).
``apply_smoother`` returns alpha * S^-1 B; the block variants solve with a partial-pivoting LU
(LAPACK getrf / getrs) per element, exactly the reference's ``block \\ B[inds, j]``.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def _dense(B):
    if sp.issparse(B):
        return B.toarray()          # the scripts pass sparse A (tests/dg_smoother_test.jl:105)
    return np.asarray(B, dtype=np.float64)


class JacobiSmoother:
    def __init__(self, mJac):
        self.mJac = np.asarray(mJac, dtype=np.float64)


class _BlockSmoother:
    def __init__(self, mBlocks, mBlockInds):
        self.mBlocks = mBlocks              # list of (lu, piv) from scipy.linalg.lu_factor
        self.mBlockInds = mBlockInds        # (p+1, n) 0-based, one column per block


class BlockJacobi(_BlockSmoother):
    pass


class AdditiveSchwarzSmoother(_BlockSmoother):
    pass


class HybridSchwarzSmoother(_BlockSmoother):
    def __init__(self, mBlocks, mBlockInds, mCountingMatrix):
        super().__init__(mBlocks, mBlockInds)
        self.mCountingMatrix = mCountingMatrix


def apply_smoother(A, B, alpha=1.0):
    B = _dense(B)
    if isinstance(A, JacobiSmoother):
        if B.ndim == 1:
            return alpha * (B / A.mJac)
        return alpha * (B / A.mJac[:, None])
    vec = B.ndim == 1
    B2 = B.reshape(B.shape[0], -1)
    Y = np.zeros(B2.shape)
    if isinstance(A, HybridSchwarzSmoother):
        for j in range(B2.shape[1]):
            temp = np.zeros(B2.shape[0])
            for i, block in enumerate(A.mBlocks):
                idx = A.mBlockInds[:, i]
                temp[idx] += sla.lu_solve(block, B2[idx, j])
            Y[:, j] = temp / A.mCountingMatrix
    else:
        for j in range(B2.shape[1]):
            for i, block in enumerate(A.mBlocks):
                idx = A.mBlockInds[:, i]
                Y[idx, j] += sla.lu_solve(block, B2[idx, j])
    Y = alpha * Y
    return Y[:, 0] if vec else Y


def _element_blocks(mesh, A):
    n = len(mesh.mElements)
    p = mesh.mP
    Ac = sp.csc_matrix(A)
    blocks = [None] * n
    inds = np.zeros((p + 1, n), dtype=np.int64)
    for i, el in enumerate(mesh.mElements):
        idx = np.asarray(el.mNodesInd)
        blocks[i] = sla.lu_factor(Ac[idx, :][:, idx].toarray())
        inds[:, i] = idx
    return blocks, inds


def cg_smoother(cgMesh, A, smootherType):
    if smootherType == "jac":
        return JacobiSmoother(sp.csc_matrix(A).diagonal().copy())
    if smootherType == "addSchwarz":
        return AdditiveSchwarzSmoother(*_element_blocks(cgMesh, A))
    if smootherType == "hybridSchwarz":
        blocks, inds = _element_blocks(cgMesh, A)
        count = np.zeros(A.shape[0])
        for el in cgMesh.mElements:
            for l in el.mNodesInd:
                count[l] += 1.0
        return HybridSchwarzSmoother(blocks, inds, count)
    raise ValueError(smootherType)


def dg_smoother(dgMesh, A, smootherType):
    if smootherType == "jac":
        return JacobiSmoother(sp.csc_matrix(A).diagonal().copy())
    if smootherType == "blockJac":
        return BlockJacobi(*_element_blocks(dgMesh, A))
    raise ValueError(smootherType)
