"""Host set-up: continuous-Galerkin discretisation of -u'' = f (vectorised assembly).

Mirrors src/cg_mesh.jl:26-48 (``CgElement``), :54-80 (``CgMesh``), :87-122 (``cg_stiffness``),
:125-185 (``cg_stiffness_and_rhs``), :188-247 (``cg_rhs``).  DOF numbering (0-based): vertices
0..n keep their mesh index, element k's p-1 interior nodes follow consecutively
(src/cg_mesh.jl:35-45, :59-63).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .dg_mesh import ElementView, _Elements, eval_func
from .reference_element import ReferenceElement


class CgMesh:
    def __init__(self, mesh, mP):
        if mP < 1:
            raise ValueError("CG needs p >= 1")
        self.mP = int(mP)
        self.mRefEl = ReferenceElement(self.mP)
        n = mesh.nFaces
        p = self.mP
        xl, xr = mesh.mVertexX[:-1], mesh.mVertexX[1:]
        self.mH = xr - xl
        self.mXc = (xl + xr) / 2.0
        self.mJacobian = self.mH / 2.0
        nodes = np.zeros((n, p + 1), dtype=np.int64)
        nodes[:, 0] = np.arange(n)
        nodes[:, 1] = np.arange(1, n + 1)
        if p > 1:
            nodes[:, 2:] = (n + 1) + np.arange(n)[:, None] * (p - 1) + np.arange(p - 1)[None, :]
        self.mNodesInd = nodes
        self.mNumNodes = (n + 1) + n * (p - 1)
        self.mNodesX = self.mXc[:, None] + (self.mH / 2.0)[:, None] * self.mRefEl.mNodesX[None, :]
        self.mNodesX[:, 0] = xl          # vertices keep their exact mesh coordinate (:38-41)
        self.mNodesX[:, 1] = xr
        rows = np.repeat(nodes[:, :, None], p + 1, axis=2).ravel()
        cols = np.repeat(nodes[:, None, :], p + 1, axis=1).ravel()
        vals = (self.mJacobian[:, None, None] * self.mRefEl.mMassMatrix[None]).ravel()
        self.mMassMatrix = sp.csc_matrix((vals, (rows, cols)), shape=(self.mNumNodes, self.mNumNodes))
        self._lu = None
        self.mElements = _Elements(self)

    @property
    def mMassMatrixLU(self):
        if self._lu is None:
            self._lu = spla.splu(self.mMassMatrix)
        return self._lu

    def _element(self, k):
        return ElementView(self, k)


def _stiffness(cgMesh):
    refEl = cgMesh.mRefEl
    nodes = cgMesh.mNodesInd
    n, m = nodes.shape
    kref = np.einsum("l,li,lj->ij", refEl.mGaussQuadWeights, refEl.mBasisGQDerivVal,
                     refEl.mBasisGQDerivVal)
    rows = np.repeat(nodes[:, :, None], m, axis=2).ravel()
    cols = np.repeat(nodes[:, None, :], m, axis=1).ravel()
    vals = ((1.0 / cgMesh.mJacobian)[:, None, None] * kref[None]).ravel()
    return sp.csc_matrix((vals, (rows, cols)), shape=(cgMesh.mNumNodes, cgMesh.mNumNodes))


def _strong_dirichlet(A, dirNodes):
    """Rows and columns of Dirichlet DOFs zeroed, 1 on their diagonal (:116-119, :179-182)."""
    if len(dirNodes) == 0:
        return A.tocsc()
    N = A.shape[0]
    keep = np.ones(N)
    keep[dirNodes] = 0.0
    Dk = sp.diags(keep)
    A = (Dk @ A @ Dk).tocsc()
    A = A + sp.csc_matrix((np.ones(len(dirNodes)), (dirNodes, dirNodes)), shape=(N, N))
    A.eliminate_zeros()
    return A.tocsc()


def cg_stiffness(cgMesh, bdCond):
    return _strong_dirichlet(_stiffness(cgMesh), np.asarray(bdCond.mDirNodes, dtype=np.int64))


def _volume_rhs(cgMesh, func):
    refEl = cgMesh.mRefEl
    xq = cgMesh.mXc[:, None] + (cgMesh.mH / 2.0)[:, None] * refEl.mGaussQuadNodes[None, :]
    fq = eval_func(func, xq)
    fe = cgMesh.mJacobian[:, None] * np.einsum("l,li,nl->ni", refEl.mGaussQuadWeights,
                                               refEl.mBasisGQFunVal, fq)
    f = np.zeros(cgMesh.mNumNodes)
    np.add.at(f, cgMesh.mNodesInd.ravel(), fe.ravel())
    return f


def _neumann_rhs(f, mesh, bdCond):
    for v in bdCond.mNeuNodes:
        side = 0 if v == 0 else 1
        f[v] += (-1.0 if side == 0 else 1.0) * bdCond.value(side)


def cg_stiffness_and_rhs(cgMesh, mesh, func, bdCond):
    A = _stiffness(cgMesh)
    f = _volume_rhs(cgMesh, func)
    _neumann_rhs(f, mesh, bdCond)
    dirNodes = np.asarray(bdCond.mDirNodes, dtype=np.int64)
    if len(dirNodes):
        f += -(A[:, dirNodes] @ np.asarray(bdCond.mDirVals, dtype=np.float64))
        f[dirNodes] = bdCond.mDirVals
    return _strong_dirichlet(A, dirNodes), f


def cg_rhs(cgMesh, mesh, func, bdCond):
    A = _stiffness(cgMesh)
    f = _volume_rhs(cgMesh, func)
    dirNodes = np.asarray(bdCond.mDirNodes, dtype=np.int64)
    if len(dirNodes):
        f += -(A[:, dirNodes] @ np.asarray(bdCond.mDirVals, dtype=np.float64))
    _neumann_rhs(f, mesh, bdCond)
    if len(dirNodes):
        f[dirNodes] = bdCond.mDirVals
    return f
