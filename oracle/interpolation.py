"""Oracle (test infrastructure): prolongation matrices L (fine rows x coarse columns).

Follows src/interpolation.jl:5-55 (``cg_cg``), :91-109 (``dg_dg``), :145-220 (``dg_cg``, flags
0/1/2), :226-264 (``aggdg_aggdg``), :270-292 (``aggdg_dg``), :330-410 (``aggdg_cg``, flags 0/1/2).
Restriction is always the transpose (src/solvers.jl:36).
"""
import numpy as np
import scipy.sparse as sp

from .aggdg import evaluate_local_modal_basis_fun
from .refelem import evaluate_nodal_basis_fun, gauss_quad
from .refmesh import isBoundary


def _sparse(data, shape):
    r, c, v = zip(*data)
    return sp.csc_matrix((np.array(v, dtype=np.float64), (np.array(r), np.array(c))), shape=shape)


def cg_cg_interpolation(lowMesh, highMesh):
    lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
    L = sp.lil_matrix((highMesh.mNumNodes, lowMesh.mNumNodes))
    eps = np.finfo(np.float64).eps
    for k, lowEl in enumerate(lowMesh.mElements):
        highEl = highMesh.mElements[k]
        for j in range(len(lowEl.mNodesInd)):
            lowNode = lowEl.mNodesInd[j]
            for i in range(2, len(highEl.mNodesInd)):
                highNode = highEl.mNodesInd[i]
                if abs(L[highNode, lowNode]) <= eps:
                    L[highNode, lowNode] = lowVal[i, j]
        for j in range(2):
            L[highEl.mNodesInd[j], lowEl.mNodesInd[j]] = lowVal[j, j]
    return L.tocsc()


def dg_dg_interpolation(lowMesh, highMesh):
    lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
    data = []
    for k, lowEl in enumerate(lowMesh.mElements):
        highEl = highMesh.mElements[k]
        for j, lowNode in enumerate(lowEl.mNodesInd):
            for i, highNode in enumerate(highEl.mNodesInd):
                data.append((highNode, lowNode, lowVal[i, j]))
    return _sparse(data, (highMesh.mNumNodes, lowMesh.mNumNodes))


def _lumped_solve(N, highMesh):
    """``Diagonal(rowsum(M)) \\ N`` (src/interpolation.jl:210-217): a division, not a multiply by
    the reciprocal."""
    lumped = np.zeros(highMesh.mNumNodes)
    Mr = highMesh.mMassMatrix.tocsr()
    for j in range(highMesh.mNumNodes):
        lumped[j] = Mr[j, :].sum()
    Nc = N.tocoo()
    return sp.csc_matrix((Nc.data / lumped[Nc.row], (Nc.row, Nc.col)), shape=N.shape)


def dg_cg_interpolation(lowMesh, highMesh, mesh, interpFlag):
    if interpFlag in (0, 1):
        gq, gqw = gauss_quad(lowMesh.mP + highMesh.mP)
        highGQ = evaluate_nodal_basis_fun(highMesh.mRefEl.mBasisFunCoeff, gq)
        lowGQ = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, gq)
        data = []
        for k, lowEl in enumerate(lowMesh.mElements):
            highEl = highMesh.mElements[k]
            temp = np.zeros((len(highEl.mNodesInd), len(lowEl.mNodesInd)))
            for j in range(len(lowEl.mNodesInd)):
                for i in range(len(highEl.mNodesInd)):
                    for l in range(len(gq)):
                        temp[i, j] += lowEl.mJacobian * gqw[l] * highGQ[l, i] * lowGQ[l, j]
            for j, lowNode in enumerate(lowEl.mNodesInd):
                for i, highNode in enumerate(highEl.mNodesInd):
                    data.append((highNode, lowNode, temp[i, j]))
        N = _sparse(data, (highMesh.mNumNodes, lowMesh.mNumNodes))
        if interpFlag == 0:
            return highMesh.mMassMatrixLU.solve(N.toarray())
        return _lumped_solve(N, highMesh)
    if interpFlag == 2:
        lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
        data = []
        for k, lowEl in enumerate(lowMesh.mElements):
            highEl = highMesh.mElements[k]
            for j in range(len(lowEl.mNodesInd)):
                lowNode = lowEl.mNodesInd[j]
                for i in range(2):
                    highNode = highEl.mNodesInd[i]
                    vert = mesh.mVertices[highNode]
                    w = 1.0 if isBoundary(vert) else 0.5
                    data.append((highNode, lowNode, w * lowVal[i, j]))
                for i in range(2, len(highEl.mNodesInd)):
                    data.append((highEl.mNodesInd[i], lowNode, lowVal[i, j]))
        return _sparse(data, (highMesh.mNumNodes, lowMesh.mNumNodes))
    raise ValueError("Only implemented for interpFlag = 0, 1, or 2.")


def aggdg_aggdg_interpolation(coarseMesh, fineMesh, baseMesh):
    if coarseMesh.mP != fineMesh.mP:
        raise ValueError("The two agglomerated meshes must have the same p.")
    p = coarseMesh.mP
    gqw = fineMesh.mGaussQuadWeights
    data = []
    for coarseEl in coarseMesh.mElements:
        count = 0
        for fineElInd in coarseEl.mSubAggElementInds:
            fineEl = fineMesh.mElements[fineElInd - 1]
            temp = np.zeros((p + 1, p + 1))
            for k, baseElInd in enumerate(fineEl.mBaseElementInds):
                baseEl = baseMesh.mElements[baseElInd - 1]
                for j in range(len(coarseEl.mNodesInd)):
                    for i in range(len(fineEl.mNodesInd)):
                        for l in range(len(gqw)):
                            temp[i, j] += (baseEl.mJacobian * gqw[l] * fineEl.mBasisGQFunVal[k][l, i]
                                           * coarseEl.mBasisGQFunVal[count + k][l, j])
            count += len(fineEl.mBaseElementInds)
            for j, node2 in enumerate(coarseEl.mNodesInd):
                for i, node1 in enumerate(fineEl.mNodesInd):
                    data.append((node1, node2, temp[i, j]))
    N = _sparse(data, (fineMesh.mNumNodes, coarseMesh.mNumNodes))
    return fineMesh.mMassMatrixLU.solve(N)


def aggdg_dg_interpolation(aggMesh, baseMesh):
    refEl = baseMesh.mRefEl
    data = []
    for aggEl in aggMesh.mElements:
        for baseElInd in aggEl.mBaseElementInds:
            baseEl = baseMesh.mElements[baseElInd - 1]
            val = evaluate_local_modal_basis_fun(aggMesh.mP, aggEl.mBoundingBox,
                                                 baseEl.mRefMap(refEl.mNodesX))
            for j, aggNode in enumerate(aggEl.mNodesInd):
                for i, baseNode in enumerate(baseEl.mNodesInd):
                    data.append((baseNode, aggNode, val[i, j]))
    return _sparse(data, (baseMesh.mNumNodes, aggMesh.mNumNodes))


def aggdg_cg_interpolation(aggMesh, baseMesh, mesh, interpFlag):
    refEl = baseMesh.mRefEl
    if interpFlag in (0, 1):
        gq, gqw = refEl.mGaussQuadNodes, refEl.mGaussQuadWeights
        data = []
        for aggEl in aggMesh.mElements:
            for baseElInd in aggEl.mBaseElementInds:
                baseEl = baseMesh.mElements[baseElInd - 1]
                temp = np.zeros((len(baseEl.mNodesInd), len(aggEl.mNodesInd)))
                aggGQ = evaluate_local_modal_basis_fun(aggMesh.mP, aggEl.mBoundingBox,
                                                       baseEl.mRefMap(gq))
                for j in range(len(aggEl.mNodesInd)):
                    for i in range(len(baseEl.mNodesInd)):
                        for l in range(len(gq)):
                            temp[i, j] += (baseEl.mJacobian * gqw[l] * refEl.mBasisGQFunVal[l, i]
                                           * aggGQ[l, j])
                for j, aggNode in enumerate(aggEl.mNodesInd):
                    for i, baseNode in enumerate(baseEl.mNodesInd):
                        data.append((baseNode, aggNode, temp[i, j]))
        N = _sparse(data, (baseMesh.mNumNodes, aggMesh.mNumNodes))
        if interpFlag == 0:
            return baseMesh.mMassMatrixLU.solve(N.toarray())
        return _lumped_solve(N, baseMesh)
    if interpFlag == 2:
        data = []
        for aggEl in aggMesh.mElements:
            for baseElInd in aggEl.mBaseElementInds:
                baseEl = baseMesh.mElements[baseElInd - 1]
                val = evaluate_local_modal_basis_fun(aggMesh.mP, aggEl.mBoundingBox,
                                                     baseEl.mRefMap(refEl.mNodesX))
                for j in range(len(aggEl.mNodesInd)):
                    aggNode = aggEl.mNodesInd[j]
                    for i in range(2):
                        baseNode = baseEl.mNodesInd[i]
                        vert = mesh.mVertices[baseNode]
                        w = 1.0 if isBoundary(vert) else 0.5
                        data.append((baseNode, aggNode, w * val[i, j]))
                    for i in range(2, len(baseEl.mNodesInd)):
                        data.append((baseEl.mNodesInd[i], aggNode, val[i, j]))
        return _sparse(data, (baseMesh.mNumNodes, aggMesh.mNumNodes))
    raise ValueError("Only implemented for interpFlag = 0, 1, or 2.")
