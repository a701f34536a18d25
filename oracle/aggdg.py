"""Oracle (test infrastructure): agglomerated DG meshes (modal basis on the bounding box, p in {0,1}).

Follows src/agglomerated_dg_mesh.jl:1-72 (types), :183-278 (``AgglomeratedDgElement1``),
:297-327 (modal basis), :400-495 (``AgglomeratedDgMesh1`` from an agglomeration map),
:501-559 (``AgglomeratedDgElementN``), :596-635 (``AgglomeratedDgMeshN``),
:641-873 (``dg_flux_operators``), :875-994 (``dg_flux_rhs``).
Agglomeration maps ``agg`` hold 1-based base-element ids exactly as the reference scripts build
them (tests/full_heirarchy_test.jl:63-75); DOF indices (mNodesInd) are 0-based.
"""
import numpy as np

from .block_diagonal import BlockDiagonal
from .dg import _coo, compute_switch
from .refelem import gauss_quad


class AgglomeratedDgVertex:
    def __init__(self, mIndex, mX):
        self.mIndex = mIndex
        self.mX = mX
        self.mFaces = [0, 0]


def _is_bd(v):
    return v.mFaces[1] < 1


def evaluate_local_modal_basis_fun(p, boundingBox, nodes):
    """phi_1 = 1, phi_2 = 2 (x - xc) / h on the bounding box (:297-315)."""
    nodes = np.atleast_1d(np.asarray(nodes, dtype=np.float64))
    val = np.zeros((len(nodes), p + 1))
    if p == 0:
        val[:, 0] = 1.0
        return val
    if p == 1:
        xC = (boundingBox[0] + boundingBox[1]) / 2.0
        h = boundingBox[1] - boundingBox[0]
        val[:, 0] = 1.0
        val[:, 1] = 2 * (nodes - xC) / h
        return val
    raise ValueError("Only implemented for p = 0 and p = 1.")


def evaluate_local_modal_basis_deriv(p, boundingBox):
    """:317-327."""
    if p == 0:
        return np.array([0.0])
    if p == 1:
        h = boundingBox[1] - boundingBox[0]
        return np.array([0.0, 2.0 / h])
    raise ValueError("Only implemented for p = 0 and p = 1.")


class AgglomeratedDgElement1:
    def __init__(self, mIndex, mP, mBaseElementInds, baseMesh, mesh, allVertices, gaussQuadNodes):
        self.mIndex = mIndex
        self.mP = mP
        self.mNodesInd = np.arange((mIndex - 1) * (mP + 1), mIndex * (mP + 1), dtype=np.int64)
        self.mBaseElementInds = list(mBaseElementInds)
        self.mSubAggElementInds = self.mBaseElementInds
        min_x, max_x = np.inf, -np.inf
        for elInd in self.mBaseElementInds:
            min_x = min(min_x, baseMesh.mElements[elInd - 1].mNodesX[0])
            max_x = max(max_x, baseMesh.mElements[elInd - 1].mNodesX[1])
        self.mBoundingBox = [min_x, max_x]
        self.mBasisGQFunVal = []
        for elInd in self.mBaseElementInds:
            el = baseMesh.mElements[elInd - 1]
            self.mBasisGQFunVal.append(evaluate_local_modal_basis_fun(
                mP, self.mBoundingBox, el.mRefMap(np.asarray(gaussQuadNodes))))
        self.mBasisDerivVal = evaluate_local_modal_basis_deriv(mP, self.mBoundingBox)
        self.mVertices = []
        for i in self.mBaseElementInds:
            face = mesh.mFaces[i - 1]
            for vert in face.mVertices:
                isBd = not ((vert.mFaces[0] in self.mBaseElementInds)
                            and (vert.mFaces[1] in self.mBaseElementInds))
                if isBd:
                    self.mVertices.append(allVertices[vert.mIndex - 1])
        self.mBdBasisGQFunVal = [
            evaluate_local_modal_basis_fun(mP, self.mBoundingBox, v.mX)[0, :] for v in self.mVertices]


def _mass_matrix(mP, elements, baseMesh, gqw):
    blocks = [None] * len(elements)
    inds = np.zeros((mP + 1, len(elements)), dtype=np.int64)
    for el in elements:
        n = len(el.mNodesInd)
        temp = np.zeros((n, n))
        for k, baseElInd in enumerate(el.mBaseElementInds):
            baseEl = baseMesh.mElements[baseElInd - 1]
            for j in range(n):
                for i in range(n):
                    for l in range(len(gqw)):
                        temp[i, j] += (baseEl.mJacobian * gqw[l] * el.mBasisGQFunVal[k][l, i]
                                       * el.mBasisGQFunVal[k][l, j])
        blocks[el.mIndex - 1] = temp
        inds[:, el.mIndex - 1] = el.mNodesInd
    return BlockDiagonal(blocks, mP + 1, inds)


class AgglomeratedDgMesh1:
    """First agglomerated level: unions of base (CG or DG) elements (:400-495)."""

    def __init__(self, mP, agg, mesh, baseMesh):
        self.mP = mP
        self.mGaussQuadNodes, self.mGaussQuadWeights = gauss_quad(2 * mP)
        self.mAllVertices = [AgglomeratedDgVertex(v.mIndex, v.mX) for v in mesh.mVertices]
        self.mElements = [
            AgglomeratedDgElement1(k + 1, mP, baseElInds, baseMesh, mesh, self.mAllVertices,
                                   self.mGaussQuadNodes)
            for k, baseElInds in enumerate(agg)]
        self.mNumNodes = len(self.mElements) * (mP + 1)
        self.mVertices = []
        for el in self.mElements:
            for vert in el.mVertices:
                if vert.mFaces[0] == 0:
                    vert.mFaces[0] = el.mIndex
                    self.mVertices.append(vert)
                elif vert.mFaces[1] == 0:
                    vert.mFaces[1] = el.mIndex
                else:
                    raise RuntimeError("Vertex can only neighbor two elements.")
        self.mMassMatrix = _mass_matrix(mP, self.mElements, baseMesh, self.mGaussQuadWeights)
        self.mMassMatrixLU = self.mMassMatrix.lu()
        self.mSwitch = compute_switch(self.mVertices,
                                      lambda v: self.mElements[v.mFaces[0] - 1].mVertices)


class AgglomeratedDgElementN:
    def __init__(self, mIndex, mP, mSubAggElementInds, subAggMesh, baseMesh):
        self.mIndex = mIndex
        self.mP = mP
        self.mNodesInd = np.arange((mIndex - 1) * (mP + 1), mIndex * (mP + 1), dtype=np.int64)
        self.mSubAggElementInds = list(mSubAggElementInds)
        self.mBaseElementInds = []
        for elInd in self.mSubAggElementInds:
            self.mBaseElementInds.extend(subAggMesh.mElements[elInd - 1].mBaseElementInds)
        min_x, max_x = np.inf, -np.inf
        for elInd in self.mSubAggElementInds:
            min_x = min(min_x, subAggMesh.mElements[elInd - 1].mBoundingBox[0])
            max_x = max(max_x, subAggMesh.mElements[elInd - 1].mBoundingBox[1])
        self.mBoundingBox = [min_x, max_x]
        gq, _ = gauss_quad(2 * mP)
        self.mBasisGQFunVal = []
        for elInd in self.mBaseElementInds:
            el = baseMesh.mElements[elInd - 1]
            self.mBasisGQFunVal.append(evaluate_local_modal_basis_fun(
                mP, self.mBoundingBox, el.mRefMap(np.asarray(gq))))
        self.mBasisDerivVal = evaluate_local_modal_basis_deriv(mP, self.mBoundingBox)


class AgglomeratedDgMeshN:
    """Deeper agglomerated levels: unions of elements of the previous agglomerated mesh (:596-635)."""

    def __init__(self, mP, agg, subAggMesh, baseMesh):
        self.mP = mP
        self.mGaussQuadNodes, self.mGaussQuadWeights = gauss_quad(2 * mP)
        self.mElements = [AgglomeratedDgElementN(k + 1, mP, sub, subAggMesh, baseMesh)
                          for k, sub in enumerate(agg)]
        self.mNumNodes = len(self.mElements) * (mP + 1)
        self.mMassMatrix = _mass_matrix(mP, self.mElements, baseMesh, self.mGaussQuadWeights)
        self.mMassMatrixLU = self.mMassMatrix.lu()


def _outer(data, rows, cols, scale, a, b):
    for j, node2 in enumerate(cols):
        for i, node1 in enumerate(rows):
            data.append((node1, node2, scale * a[i] * b[j]))


def agg_dg_flux_operators(aggDgMesh, baseMesh, bdCond, CDir):
    """``dg_flux_operators(::AgglomeratedDgMesh1, baseMesh, bdCond, CDir)`` (:641-873).
    The p = 0 branch of the reference is the same rank-1 formula restricted to one basis
    function (and no volume term), so both branches are written once here."""
    dG, dD, dC = [], [], []
    gqw = aggDgMesh.mGaussQuadWeights
    if aggDgMesh.mP >= 1:
        for el in aggDgMesh.mElements:
            n = len(el.mNodesInd)
            temp = np.zeros((n, n))
            for k, baseElInd in enumerate(el.mBaseElementInds):
                baseEl = baseMesh.mElements[baseElInd - 1]
                for j in range(n):
                    for i in range(n):
                        for l in range(len(gqw)):
                            temp[i, j] += (baseEl.mJacobian * gqw[l] * el.mBasisDerivVal[i]
                                           * el.mBasisGQFunVal[k][l, j])
            for j, node2 in enumerate(el.mNodesInd):
                for i, node1 in enumerate(el.mNodesInd):
                    dG.append((node1, node2, temp[i, j]))
                    dD.append((node1, node2, temp[i, j]))
    for i, vert in enumerate(aggDgMesh.mVertices):
        if _is_bd(vert):
            el = aggDgMesh.mElements[vert.mFaces[0] - 1]
            if vert is el.mVertices[0]:
                sign, side = -1.0, 0
            elif vert is el.mVertices[1]:
                sign, side = 1.0, 1
            else:
                raise RuntimeError("vertex / element mismatch")
            phi = el.mBdBasisGQFunVal[side]
            if vert.mIndex in bdCond.mDirNodes:
                _outer(dD, el.mNodesInd, el.mNodesInd, -sign, phi, phi)
                _outer(dC, el.mNodesInd, el.mNodesInd, CDir, phi, phi)
            elif vert.mIndex in bdCond.mNeuNodes:
                _outer(dG, el.mNodesInd, el.mNodesInd, -sign, phi, phi)
            else:
                raise RuntimeError("Boundary vertex is not included in the boundary condition.")
        else:
            S = aggDgMesh.mSwitch[i]
            uhatEl = aggDgMesh.mElements[vert.mFaces[S - 1] - 1]
            qhatEl = aggDgMesh.mElements[vert.mFaces[S % 2] - 1]
            for k in vert.mFaces:
                el = aggDgMesh.mElements[k - 1]
                if vert is el.mVertices[0]:
                    sign, side = -1.0, 0
                elif vert is el.mVertices[1]:
                    sign, side = 1.0, 1
                else:
                    raise RuntimeError("vertex / element mismatch")
                _outer(dG, el.mNodesInd, uhatEl.mNodesInd, -sign,
                       el.mBdBasisGQFunVal[side], uhatEl.mBdBasisGQFunVal[1])
                _outer(dD, el.mNodesInd, qhatEl.mNodesInd, -sign,
                       el.mBdBasisGQFunVal[side], qhatEl.mBdBasisGQFunVal[0])
    N = aggDgMesh.mNumNodes
    return _coo(dG, N), _coo(dD, N), _coo(dC, N)


def agg_dg_flux_rhs(aggDgMesh, baseMesh, func, bdCond, CDir):
    """``dg_flux_rhs(::AgglomeratedDgMesh1, ...)`` (:875-994)."""
    f = np.zeros(aggDgMesh.mNumNodes)
    r = np.zeros(aggDgMesh.mNumNodes)
    gq, gqw = aggDgMesh.mGaussQuadNodes, aggDgMesh.mGaussQuadWeights
    for el in aggDgMesh.mElements:
        for k, baseElInd in enumerate(el.mBaseElementInds):
            baseEl = baseMesh.mElements[baseElInd - 1]
            for i, node in enumerate(el.mNodesInd):
                for l in range(len(gq)):
                    f[node] += (baseEl.mJacobian * gqw[l] * el.mBasisGQFunVal[k][l, i]
                                * func(baseEl.mRefMap(gq[l])))
    for i, nodeIdx in enumerate(bdCond.mDirNodes):
        vert = aggDgMesh.mAllVertices[nodeIdx - 1]
        dirVal = bdCond.mDirVals[i]
        el = aggDgMesh.mElements[vert.mFaces[0] - 1]
        if vert is el.mVertices[0]:
            sign, side = -1.0, 0
        elif vert is el.mVertices[1]:
            sign, side = 1.0, 1
        else:
            raise RuntimeError("vertex / element mismatch")
        for ii, node in enumerate(el.mNodesInd):
            f[node] += CDir * dirVal * el.mBdBasisGQFunVal[side][ii]
            r[node] += sign * dirVal * el.mBdBasisGQFunVal[side][ii]
    for nodeIdx in bdCond.mNeuNodes:
        vert = aggDgMesh.mAllVertices[nodeIdx - 1]
        el = aggDgMesh.mElements[vert.mFaces[0] - 1]
        if vert is el.mVertices[0]:
            sign, side, neuVal = -1.0, 0, bdCond.mBdCond[0][1]
        elif vert is el.mVertices[1]:
            sign, side, neuVal = 1.0, 1, bdCond.mBdCond[1][1]
        else:
            raise RuntimeError("vertex / element mismatch")
        for ii, node in enumerate(el.mNodesInd):
            f[node] += sign * neuVal * el.mBdBasisGQFunVal[side][ii]
    return f, r
