"""Host set-up: nodal DG discretisation of -u'' = f with LDG-type fluxes (vectorised assembly).

Mirrors src/dg_mesh.jl:32-52 (``DgElement``: mNodesInd[i] = (k-1)(p+1)+i), :58-138 (``DgMesh``),
:144-336 (``dg_flux_operators`` -> G, D, C), :342-457 (``dg_flux_rhs`` -> f, r).  The reference's
switch computation compares a face with itself (:96-97), so at every interior vertex
uhat = u_L (left element), qhat = q_R (right element); that behaviour is reproduced, not "fixed".
All indices are 0-based (reference - 1).
"""
import numpy as np
import scipy.sparse as sp

from .block_diagonal import BlockDiagonal
from .reference_element import ReferenceElement


def eval_func(func, x):
    """Evaluate a scalar python function on an array (vectorised call when the function allows)."""
    x = np.asarray(x, dtype=np.float64)
    try:
        y = np.asarray(func(x), dtype=np.float64)
        if y.shape == x.shape:
            return y
        if y.shape == ():
            return np.full(x.shape, float(y))
    except Exception:
        pass
    return np.vectorize(func, otypes=[np.float64])(x)


class ElementView:
    """Read-only stand-in for one entry of the reference's ``mesh.mElements``."""

    def __init__(self, mesh, k):
        self.mIndex = k
        self.mP = mesh.mP
        self.mNodesInd = mesh.mNodesInd[k]
        self.mNodesX = mesh.mNodesX[k] if getattr(mesh, "mNodesX", None) is not None else None
        self.mJacobian = mesh.mJacobian[k] if getattr(mesh, "mJacobian", None) is not None else None


class _Elements:
    def __init__(self, mesh):
        self._m = mesh

    def __len__(self):
        return self._m.mNodesInd.shape[0]

    def __getitem__(self, k):
        if k < 0:
            k += len(self)
        return self._m._element(k)

    def __iter__(self):
        return (self._m._element(k) for k in range(len(self)))


class DgMesh:
    def __init__(self, mesh, mP):
        self.mP = int(mP)
        self.mRefEl = ReferenceElement(self.mP)
        m = self.mP + 1
        n = mesh.nFaces
        xl, xr = mesh.mVertexX[:-1], mesh.mVertexX[1:]
        self.mH = xr - xl
        self.mXc = (xl + xr) / 2.0
        self.mJacobian = self.mH / 2.0
        self.mNodesInd = np.arange(n * m, dtype=np.int64).reshape(n, m)
        self.mNodesX = self.mXc[:, None] + (self.mH / 2.0)[:, None] * self.mRefEl.mNodesX[None, :]
        self.mNumNodes = n * m
        self.mMassMatrix = BlockDiagonal(self.mJacobian[:, None, None] * self.mRefEl.mMassMatrix[None],
                                         m, self.mNodesInd.T)
        self.mMassMatrixLU = self.mMassMatrix.lu()
        # 1-based as in the reference: S = 1 -> uhat from vert.mFaces[1] (the left face)
        self.mSwitch = np.ones(mesh.nVertices, dtype=np.int64)
        self.mSwitch[0] = 2
        self.mElements = _Elements(self)

    def _element(self, k):
        return ElementView(self, k)

    def ref_map(self, k, xi):
        """``mRefMap`` of element(s) k: x = xc + (h/2) xi."""
        return self.mXc[k] + self.mH[k] / 2.0 * xi


def _csc(rows, cols, vals, N):
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(N, N))


def dg_flux_operators(dgMesh, mesh, bdCond, CDir):
    """G (gradient), D (divergence), C (Dirichlet penalty) as CSC (src/dg_mesh.jl:144-336)."""
    refEl = dgMesh.mRefEl
    nodes = dgMesh.mNodesInd
    n, m = nodes.shape
    N = dgMesh.mNumNodes
    e1 = 0
    e2 = 1 if dgMesh.mP >= 1 else 0
    gr, gc, gv = [], [], []
    dr, dc, dv = [], [], []
    cr, cc, cv = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], [np.zeros(0)]
    if dgMesh.mP >= 1:
        vol = np.einsum("l,li,lj->ij", refEl.mGaussQuadWeights, refEl.mBasisGQDerivVal,
                        refEl.mBasisGQFunVal)
        rows = np.repeat(nodes[:, :, None], m, axis=2).ravel()
        cols = np.repeat(nodes[:, None, :], m, axis=1).ravel()
        vals = np.tile(vol.ravel(), n)
        gr.append(rows); gc.append(cols); gv.append(vals)
        dr.append(rows); dc.append(cols); dv.append(vals)
    if n > 1:
        a2 = nodes[:-1, e2]     # right end of the left element a (uhat = u_L lives here)
        a1 = nodes[:-1, e1]
        b1 = nodes[1:, e1]      # left end of the right element b (qhat = q_R lives here)
        ua = a2 if dgMesh.mP >= 1 else a1
        one = np.ones(n - 1)
        # vertex is the outgoing boundary of a: G[a2, uhat] -= 1, D[a2, qhat] -= 1
        gr.append(a2); gc.append(ua); gv.append(-one)
        dr.append(a2); dc.append(b1); dv.append(-one)
        # vertex is the incoming boundary of b: G[b1, uhat] += 1, D[b1, qhat] += 1
        gr.append(b1); gc.append(ua); gv.append(one)
        dr.append(b1); dc.append(b1); dv.append(one)
    for side, el, loc, sgn in ((0, 0, e1, 1.0), (1, n - 1, e2, -1.0)):
        node = np.array([nodes[el, loc]])
        if mesh.mBdSide[0 if side == 0 else -1] == 0:
            raise ValueError("Boundary vertex is not included in the boundary condition.")
        if bdCond.kind(side) == "dir":
            dr.append(node); dc.append(node); dv.append(np.array([sgn]))
            cr.append(node); cc.append(node); cv.append(np.array([float(CDir)]))
        else:
            gr.append(node); gc.append(node); gv.append(np.array([sgn]))
    return _csc(gr, gc, gv, N), _csc(dr, dc, dv, N), _csc(cr, cc, cv, N)


def dg_flux_rhs(dgMesh, mesh, func, bdCond, CDir):
    """f and r of b = f - D (M \\ r) (src/dg_mesh.jl:342-457)."""
    refEl = dgMesh.mRefEl
    nodes = dgMesh.mNodesInd
    n, m = nodes.shape
    e1 = 0
    e2 = 1 if dgMesh.mP >= 1 else 0
    xq = dgMesh.mXc[:, None] + (dgMesh.mH / 2.0)[:, None] * refEl.mGaussQuadNodes[None, :]
    fq = eval_func(func, xq)                                             # (n, nq)
    f = (dgMesh.mJacobian[:, None]
         * np.einsum("l,li,nl->ni", refEl.mGaussQuadWeights, refEl.mBasisGQFunVal, fq)).ravel()
    r = np.zeros(dgMesh.mNumNodes)
    for side, el, loc, sgn in ((0, 0, e1, -1.0), (1, n - 1, e2, 1.0)):
        node = nodes[el, loc]
        val = bdCond.value(side)
        if bdCond.kind(side) == "dir":
            f[node] += CDir * val
            r[node] += sgn * val
        else:
            f[node] += sgn * val
    return f, r
