"""Smoother operator types and their constructors (host mirror of src/smoother.jl).

``JacobiSmoother`` (src/smoother.jl:52-58) and ``BlockJacobi`` (:64-81) are the two smoothers a
MeshHierarchy wires in (src/mesh_heirarchy.jl:51-176); ``cg_smoother(:jac)`` (:88-102) and
``dg_smoother(:jac | :blockJac)`` (:142-168) build them from a level operator.  Applying a smoother
is a GPU operation: ``apply_smoother`` uploads the (operator, smoother) pair once as a one-level
device hierarchy and calls amg1d_apply_smoother - there is no CPU path.

The Schwarz smoothers (:1-46, :104-135) are never used by a hierarchy; they are listed as "next" in
SURVEY 8f and raise NotImplementedError here.
"""
import numpy as np
import scipy.sparse as sp

from . import blocks as blk


class AbstractSmoother:
    _device = None     # one-level DeviceHierarchy, created on first use
    _device_A = None   # the operator that one-level hierarchy was built for
    _owner = None      # (DeviceHierarchy, level) when the smoother belongs to a MeshHierarchy


class JacobiSmoother(AbstractSmoother):
    """mJac: the diagonal of the level operator (the reference holds a ``Diagonal``)."""

    def __init__(self, mJac, A=None, slots=None):
        self.mJac = np.asarray(mJac, dtype=np.float64)
        self._A = A
        self._slots = slots


class BlockJacobi(AbstractSmoother):
    """mBlocks: (n, m, m) diagonal blocks A[el.mNodesInd, el.mNodesInd]; mBlockInds (m, n)."""

    def __init__(self, mBlocks, mBlockInds, A=None, slots=None):
        self.mBlocks = np.asarray(mBlocks, dtype=np.float64)
        self.mBlockInds = np.asarray(mBlockInds, dtype=np.int64)
        self._A = A
        self._slots = slots


def _diag_blocks(A, nodes):
    A = sp.csr_matrix(A)
    n, m = nodes.shape
    out = np.zeros((n, m, m))
    coo = A.tocoo()
    N = A.shape[0]
    elem_of = np.full(N, -1, dtype=np.int64)
    local_of = np.zeros(N, dtype=np.int64)
    elem_of[nodes.ravel()] = np.repeat(np.arange(n), m)
    local_of[nodes.ravel()] = np.tile(np.arange(m), n)
    same = (elem_of[coo.row] == elem_of[coo.col]) & (elem_of[coo.row] >= 0)
    np.add.at(out, (elem_of[coo.row[same]], local_of[coo.row[same]], local_of[coo.col[same]]),
              coo.data[same])
    return out


def cg_smoother(cgMesh, A, smootherType):
    if smootherType == "jac":
        return JacobiSmoother(sp.csc_matrix(A).diagonal().copy(), A, blk.level_slots(cgMesh))
    if smootherType in ("addSchwarz", "hybridSchwarz"):
        raise NotImplementedError("overlapping Schwarz smoothers are not part of the GPU V-cycle path")
    raise ValueError(f"unknown smoother type {smootherType!r}")


def dg_smoother(dgMesh, A, smootherType):
    slots = blk.level_slots(dgMesh)
    if smootherType == "jac":
        return JacobiSmoother(sp.csc_matrix(A).diagonal().copy(), A, slots)
    if smootherType == "blockJac":
        return BlockJacobi(_diag_blocks(A, dgMesh.mNodesInd), dgMesh.mNodesInd.T, A, slots)
    raise ValueError(f"unknown smoother type {smootherType!r}")


def smoother_inverse(smoother, slots):
    """(Dinv array in ABI layout, is_diagonal) for a smoother on a level with the given slot map."""
    ne, m = slots.shape
    if isinstance(smoother, JacobiSmoother):
        dinv = np.ones((ne, m))
        valid = slots >= 0
        dinv[valid] = 1.0 / smoother.mJac[slots[valid]]
        return np.ascontiguousarray(dinv), True
    if isinstance(smoother, BlockJacobi):
        if smoother.mBlocks.shape[0] != ne or smoother.mBlocks.shape[1] != m:
            raise ValueError("block-Jacobi blocks do not match the level's element grouping")
        return blk.to_abi(np.linalg.inv(smoother.mBlocks)), False
    raise TypeError("unsupported smoother type")
