"""Host set-up: agglomerated DG meshes (modal basis 1, 2(x-xc)/h on the bounding box; p in {0, 1}).

Mirrors src/agglomerated_dg_mesh.jl:183-278 (``AgglomeratedDgElement1``), :297-327 (modal basis),
:400-495 (``AgglomeratedDgMesh1``), :501-559 / :596-635 (``AgglomeratedDgElementN`` / ``MeshN``),
:641-873 (``dg_flux_operators``), :875-994 (``dg_flux_rhs``).

The reference has no agglomeration algorithm: its scripts hand-build contiguous index ranges
(tests/full_heirarchy_test.jl:63-75).  Accordingly an agglomeration map here is a list of contiguous,
ascending, gap-free 0-based index ranges (ragged sizes allowed); it is stored as offsets ``mStarts``
so that element K owns base elements mStarts[K] .. mStarts[K+1]-1, which is what
``mBaseElementInds`` / ``mSubAggElementInds`` contain in the reference.
"""
import numpy as np
import scipy.sparse as sp

from .block_diagonal import BlockDiagonal
from .dg_mesh import ElementView, _Elements, eval_func
from .reference_element import gauss_quad


def evaluate_local_modal_basis_fun(p, boundingBox, nodes):
    """(len(nodes), p+1) values of 1, 2(x - xc)/h (src/agglomerated_dg_mesh.jl:297-315)."""
    nodes = np.atleast_1d(np.asarray(nodes, dtype=np.float64))
    val = np.zeros((len(nodes), p + 1))
    if p == 0:
        val[:, 0] = 1.0
    elif p == 1:
        xC = (boundingBox[0] + boundingBox[1]) / 2.0
        h = boundingBox[1] - boundingBox[0]
        val[:, 0] = 1.0
        val[:, 1] = 2 * (nodes - xC) / h
    else:
        raise ValueError("Only implemented for p = 0 and p = 1.")
    return val


def evaluate_local_modal_basis_deriv(p, boundingBox):
    """src/agglomerated_dg_mesh.jl:317-327."""
    if p == 0:
        return np.array([0.0])
    if p == 1:
        return np.array([0.0, 2.0 / (boundingBox[1] - boundingBox[0])])
    raise ValueError("Only implemented for p = 0 and p = 1.")


def agglomeration_starts(agg, n_sub):
    """Validate an agglomeration map and return its offsets (length len(agg) + 1)."""
    if isinstance(agg, np.ndarray) and agg.ndim == 1:          # already offsets
        starts = agg.astype(np.int64)
    else:
        sizes = np.fromiter((len(a) for a in agg), dtype=np.int64, count=len(agg))
        starts = np.concatenate(([0], np.cumsum(sizes)))
        flat = np.concatenate([np.asarray(a, dtype=np.int64) for a in agg])
        if not np.array_equal(flat, np.arange(len(flat), dtype=np.int64)):
            raise ValueError("agglomerates must be contiguous, ascending, gap-free 0-based ranges")
    if starts[0] != 0 or starts[-1] != n_sub or np.any(np.diff(starts) < 1):
        raise ValueError("agglomeration map does not cover the sub-elements exactly once")
    return starts


def uniform_agglomeration(n_sub, factor):
    """Offsets of the map agg[j] = (factor*j):(factor*(j+1)) the reference scripts build."""
    if n_sub % factor:
        raise ValueError("agglomeration factor does not divide the element count")
    return np.arange(0, n_sub + 1, factor, dtype=np.int64)


class _AggBase:
    """Common part of AgglomeratedDgMesh1 / AgglomeratedDgMeshN."""

    def _build(self, mP, base_starts, bbox, baseMesh, ends=None):
        if mP not in (0, 1):
            raise ValueError("Only implemented for p = 0 and p = 1.")
        self.mP = int(mP)
        m = self.mP + 1
        self.mBaseMesh = baseMesh
        self.mBaseStarts = base_starts                      # base elements of element K
        nE = len(base_starts) - 1
        self.mBoundingBox = bbox                            # (nE, 2)
        self.mH = bbox[:, 1] - bbox[:, 0]
        self.mXc = (bbox[:, 0] + bbox[:, 1]) / 2.0
        self.mNodesInd = np.arange(nE * m, dtype=np.int64).reshape(nE, m)
        self.mNumNodes = nE * m
        self.mGaussQuadNodes, self.mGaussQuadWeights = gauss_quad(2 * self.mP)
        n_base = base_starts[-1]
        self.mAggOfBase = np.repeat(np.arange(nE, dtype=np.int64), np.diff(base_starts))
        # modal basis at the mapped Gauss points of every base element: (n_base, nq, m)
        bxc, bh = baseMesh.mXc[:n_base], baseMesh.mH[:n_base]
        xq = bxc[:, None] + (bh / 2.0)[:, None] * self.mGaussQuadNodes[None, :]
        self.mBasisGQFunVal = self._basis(xq, self.mAggOfBase[:, None])
        self.mBasisDerivVal = np.zeros((nE, m))
        if self.mP == 1:
            self.mBasisDerivVal[:, 1] = 2.0 / self.mH
        # basis at the two boundary vertices of each element: (nE, 2, m).  The reference evaluates
        # at the mesh vertex coordinate (vert.mX, :254-259), not at the bounding-box corner.
        if ends is None:
            ends = np.stack([bbox[:, 0], bbox[:, 1]], axis=1)
        self.mBdBasisGQFunVal = self._basis(ends, np.arange(nE)[:, None])
        w = self.mGaussQuadWeights
        per_base = baseMesh.mJacobian[:n_base, None, None] * np.einsum(
            "l,bli,blj->bij", w, self.mBasisGQFunVal, self.mBasisGQFunVal)
        blocks = np.add.reduceat(per_base, base_starts[:-1], axis=0)
        self.mMassMatrix = BlockDiagonal(blocks, m, self.mNodesInd.T)
        self.mMassMatrixLU = self.mMassMatrix.lu()
        self.mElements = _Elements(self)

    def _basis(self, x, K):
        """modal basis of element K (broadcast against x) at points x -> x.shape + (m,)"""
        out = np.zeros(x.shape + (self.mP + 1,))
        out[..., 0] = 1.0
        if self.mP == 1:
            out[..., 1] = 2 * (x - self.mXc[K]) / self.mH[K]
        return out

    def _element(self, k):
        el = ElementView.__new__(ElementView)
        el.mIndex = k
        el.mP = self.mP
        el.mNodesInd = self.mNodesInd[k]
        el.mBoundingBox = self.mBoundingBox[k]
        el.mBaseElementInds = np.arange(self.mBaseStarts[k], self.mBaseStarts[k + 1])
        el.mSubAggElementInds = np.arange(self.mSubStarts[k], self.mSubStarts[k + 1])
        return el


class AgglomeratedDgMesh1(_AggBase):
    """First agglomerated level: unions of base (CG or DG) elements
    (``AgglomeratedDgMesh1(mP, agg, mesh, baseMesh)``, src/agglomerated_dg_mesh.jl:400-495)."""

    def __init__(self, mP, agg, mesh, baseMesh):
        starts = agglomeration_starts(agg, mesh.nFaces)
        self.mSubStarts = starts
        xv = mesh.mVertexX
        # bounding box from the base elements' end nodes mNodesX[1], mNodesX[2] (:190-196)
        bbox = np.stack([baseMesh.mNodesX[starts[:-1], 0], baseMesh.mNodesX[starts[1:] - 1, 1]], axis=1)
        ends = np.stack([xv[starts[:-1]], xv[starts[1:]]], axis=1)
        self._build(mP, starts, bbox, baseMesh, ends)
        self.mVertexIds = starts                     # base-mesh vertex ids of the agglomerated vertices
        self.mBdSide = mesh.mBdSide[starts]
        self.mSwitch = np.ones(len(starts), dtype=np.int64)
        self.mSwitch[0] = 2


class AgglomeratedDgMeshN(_AggBase):
    """Deeper level: unions of elements of the previous agglomerated mesh
    (``AgglomeratedDgMeshN(mP, agg, subAggMesh, baseMesh)``, :596-635)."""

    def __init__(self, mP, agg, subAggMesh, baseMesh):
        sub = agglomeration_starts(agg, len(subAggMesh.mBaseStarts) - 1)
        self.mSubStarts = sub
        base_starts = subAggMesh.mBaseStarts[sub]
        sb = subAggMesh.mBoundingBox
        bbox = np.stack([np.minimum.reduceat(sb[:, 0], sub[:-1]),
                         np.maximum.reduceat(sb[:, 1], sub[:-1])], axis=1)
        self._build(mP, base_starts, bbox, baseMesh)


def _blocks_to_csc(entries, N):
    """entries: list of (row_nodes (k, mi), col_nodes (k, mj), vals (k, mi, mj))."""
    rows, cols, vals = [], [], []
    for rn, cn, v in entries:
        mi, mj = rn.shape[1], cn.shape[1]
        rows.append(np.repeat(rn[:, :, None], mj, axis=2).ravel())
        cols.append(np.repeat(cn[:, None, :], mi, axis=1).ravel())
        vals.append(np.asarray(v, dtype=np.float64).ravel())
    if not rows:
        return sp.csc_matrix((N, N))
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(N, N))


def agg_dg_flux_operators(aggDgMesh, baseMesh, bdCond, CDir):
    """``dg_flux_operators(::AgglomeratedDgMesh1, baseMesh, bdCond, CDir)`` (:641-873)."""
    nd = aggDgMesh.mNodesInd
    nE, m = nd.shape
    N = aggDgMesh.mNumNodes
    Lb = aggDgMesh.mBdBasisGQFunVal[:, 0, :]        # basis at each element's left end
    Rb = aggDgMesh.mBdBasisGQFunVal[:, 1, :]        # ... right end
    G, D, Cm = [], [], []
    if aggDgMesh.mP >= 1:
        n_base = aggDgMesh.mBaseStarts[-1]
        per_base = baseMesh.mJacobian[:n_base, None, None] * np.einsum(
            "l,bi,blj->bij", aggDgMesh.mGaussQuadWeights,
            aggDgMesh.mBasisDerivVal[aggDgMesh.mAggOfBase], aggDgMesh.mBasisGQFunVal)
        vol = np.add.reduceat(per_base, aggDgMesh.mBaseStarts[:-1], axis=0)
        G.append((nd, nd, vol))
        D.append((nd, nd, vol))
    if nE > 1:
        a, b = slice(0, nE - 1), slice(1, nE)      # left element a = uhat, right element b = qhat
        G.append((nd[a], nd[a], -Rb[a][:, :, None] * Rb[a][:, None, :]))
        D.append((nd[a], nd[b], -Rb[a][:, :, None] * Lb[b][:, None, :]))
        G.append((nd[b], nd[a], Lb[b][:, :, None] * Rb[a][:, None, :]))
        D.append((nd[b], nd[b], Lb[b][:, :, None] * Lb[b][:, None, :]))
    for side, el, phi, sgn in ((0, 0, Lb[0], 1.0), (1, nE - 1, Rb[nE - 1], -1.0)):
        rn = nd[el:el + 1]
        outer = (phi[:, None] * phi[None, :])[None]
        if bdCond.kind(side) == "dir":
            D.append((rn, rn, sgn * outer))
            Cm.append((rn, rn, CDir * outer))
        else:
            G.append((rn, rn, sgn * outer))
    return _blocks_to_csc(G, N), _blocks_to_csc(D, N), _blocks_to_csc(Cm, N)


def agg_dg_flux_rhs(aggDgMesh, baseMesh, func, bdCond, CDir):
    """``dg_flux_rhs(::AgglomeratedDgMesh1, ...)`` (:875-994)."""
    nd = aggDgMesh.mNodesInd
    nE, m = nd.shape
    n_base = aggDgMesh.mBaseStarts[-1]
    gq, gqw = aggDgMesh.mGaussQuadNodes, aggDgMesh.mGaussQuadWeights
    xq = baseMesh.mXc[:n_base, None] + (baseMesh.mH[:n_base] / 2.0)[:, None] * gq[None, :]
    fq = eval_func(func, xq)
    per_base = baseMesh.mJacobian[:n_base, None] * np.einsum(
        "l,bli,bl->bi", gqw, aggDgMesh.mBasisGQFunVal, fq)
    f = np.add.reduceat(per_base, aggDgMesh.mBaseStarts[:-1], axis=0).ravel()
    r = np.zeros(aggDgMesh.mNumNodes)
    for side, el, loc, sgn in ((0, 0, 0, -1.0), (1, nE - 1, 1, 1.0)):
        phi = aggDgMesh.mBdBasisGQFunVal[el, loc, :]
        val = bdCond.value(side)
        if bdCond.kind(side) == "dir":
            f[nd[el]] += CDir * val * phi
            r[nd[el]] += sgn * val * phi
        else:
            f[nd[el]] += sgn * val * phi
    return f, r
