"""Smoother operator types and their constructors (host mirror of src/smoother.jl).

``JacobiSmoother`` (src/smoother.jl:52-58) and ``BlockJacobi`` (:64-81) are the two smoothers a
MeshHierarchy wires in (src/mesh_heirarchy.jl:51-176); ``cg_smoother(:jac)`` (:88-102) and
``dg_smoother(:jac | :blockJac)`` (:142-168) build them from a level operator.  Applying a smoother
is a GPU operation: ``apply_smoother`` uploads the (operator, smoother) pair once as a one-level
device hierarchy and calls amg1d_apply_smoother - there is no CPU path.

The overlapping Schwarz smoothers of CG levels (``AdditiveSchwarzSmoother``, ``HybridSchwarzSmoother``,
:1-46, built by ``cg_smoother(:addSchwarz / :hybridSchwarz)``, :104-135; no hierarchy uses them, the
script tests/cg_smoother_test.jl does) are block TRIDIAGONAL operators in the [vertex_k, interior_k]
grouping of a CG level: the local solve of element k touches group k and the vertex slot of group k+1.
``schwarz_tridiag_blocks`` assembles that operator from the inverses of the element matrices; the
library applies it with amg1d_set_level_smoother.
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


class DeviceBlockJacobi(AbstractSmoother):
    """Block-Jacobi smoother of a level whose blocks were extracted and inverted on the GPU
    (MeshHierarchy(..., device_setup=True)); it exists only as (device hierarchy, level)."""

    def __init__(self, dev, level):
        self._owner = (dev, level)
        self._A = None
        self._slots = None


class DeviceJacobi(JacobiSmoother):
    """Point-Jacobi smoother of a level whose operator is a Galerkin product formed on the GPU
    (MeshHierarchy(..., device_setup=True), CG levels); mJac is downloaded on first use."""

    def __init__(self, dev, level, slots, n_dof):
        self._owner = (dev, level)
        self._A = None
        self._slots = slots
        self._n_dof = int(n_dof)
        self._jac = None

    @property
    def mJac(self):
        if self._jac is None:
            dev, level = self._owner
            ne, m = self._slots.shape
            _, di, _, _ = dev.get_level(level, ne, m, diag=True)
            d = np.einsum("eii->ei", di)
            self._jac = np.zeros(self._n_dof)
            valid = self._slots >= 0
            self._jac[self._slots[valid]] = d[valid]
        return self._jac


class AdditiveSchwarzSmoother(AbstractSmoother):
    """mBlocks: (n, p+1, p+1) element matrices A[el.mNodesInd, el.mNodesInd] (the reference keeps their
    LU factors); mBlockInds (p+1, n) = el.mNodesInd columns (left vertex, right vertex, interior nodes)."""

    def __init__(self, mBlocks, mBlockInds, A=None, slots=None):
        self.mBlocks = np.asarray(mBlocks, dtype=np.float64)
        self.mBlockInds = np.asarray(mBlockInds, dtype=np.int64)
        self._A = A
        self._slots = slots


class HybridSchwarzSmoother(AdditiveSchwarzSmoother):
    """As AdditiveSchwarzSmoother, with mCountingMatrix[l] = number of elements that own DOF l
    (src/smoother.jl:24-46: the summed local solves are divided by it)."""

    def __init__(self, mBlocks, mBlockInds, mCountingMatrix, A=None, slots=None):
        super().__init__(mBlocks, mBlockInds, A, slots)
        self.mCountingMatrix = np.asarray(mCountingMatrix, dtype=np.float64)


def _element_matrices(A, nodes):
    """(n, k, k) dense blocks A[nodes[e], nodes[e]] for (possibly overlapping) index sets nodes (n, k)."""
    A = sp.csr_matrix(A)
    n, k = nodes.shape
    rows = np.repeat(nodes[:, :, None], k, axis=2).ravel()
    cols = np.repeat(nodes[:, None, :], k, axis=1).ravel()
    return np.asarray(A[rows, cols]).reshape(n, k, k)


def schwarz_tridiag_blocks(smoother, slots):
    """S_lo, S_di, S_up, each (n + 1, p, p) in (e, i, j) order: the Schwarz smoother as a block-tridiagonal
    operator on the groups [vertex_k, interior nodes of element k] of a CG level (last group: closing
    vertex + padding).  Element k's local inverse B_k (local order: v_k, v_k+1, interior) contributes
        rows/cols (v_k, int_k) x (v_k, int_k)  -> S_di[k]        (v_k, int_k) x v_k+1 -> S_up[k][:, 0]
        v_k+1 x (v_k, int_k) -> S_lo[k+1][0, :]                  v_k+1 x v_k+1        -> S_di[k+1][0, 0]
    and the hybrid variant divides every row by its DOF's element count."""
    Binv = np.linalg.inv(smoother.mBlocks)                      # (n, p+1, p+1)
    n, p1, _ = Binv.shape
    p = p1 - 1
    if slots.shape != (n + 1, p):
        raise ValueError("Schwarz smoother does not match the CG level's grouping")
    own = np.array([0] + list(range(2, p1)))
    S_lo, S_di, S_up = (np.zeros((n + 1, p, p)) for _ in range(3))
    S_di[:n] = Binv[:, own][:, :, own]
    S_di[1:, 0, 0] += Binv[:, 1, 1]
    S_up[:n, :, 0] = Binv[:, own, 1]
    S_lo[1:, 0, :] = Binv[:, 1, own]
    if isinstance(smoother, HybridSchwarzSmoother):
        cnt = np.ones((n + 1, p))
        valid = slots >= 0
        cnt[valid] = smoother.mCountingMatrix[slots[valid]]
        for S in (S_lo, S_di, S_up):
            S /= cnt[:, :, None]
    return S_lo, S_di, S_up


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
        nodes = np.asarray(cgMesh.mNodesInd, dtype=np.int64)
        blocks = _element_matrices(A, nodes)
        if smootherType == "addSchwarz":
            return AdditiveSchwarzSmoother(blocks, nodes.T, A, blk.level_slots(cgMesh))
        count = np.bincount(nodes.ravel(), minlength=sp.csc_matrix(A).shape[0]).astype(np.float64)
        return HybridSchwarzSmoother(blocks, nodes.T, count, A, blk.level_slots(cgMesh))
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
    if isinstance(smoother, AdditiveSchwarzSmoother):
        # placeholder diagonal; the real operator goes up through amg1d_set_level_smoother
        return np.ones((ne, m)), True
    if isinstance(smoother, BlockJacobi):
        if smoother.mBlocks.shape[0] != ne or smoother.mBlocks.shape[1] != m:
            raise ValueError("block-Jacobi blocks do not match the level's element grouping")
        return blk.to_abi(np.linalg.inv(smoother.mBlocks)), False
    raise TypeError("unsupported smoother type")
