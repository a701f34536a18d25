"""Host set-up: prolongation matrices L (fine rows x coarse columns) as CSC, built blockwise.

Mirrors src/interpolation.jl:5-55 (``cg_cg_interpolation``), :91-109 (``dg_dg``), :145-220
(``dg_cg``, interpFlag 0 / 1 / 2), :226-264 (``aggdg_aggdg`` = M_f^-1 N), :270-292 (``aggdg_dg`` =
modal basis at the base nodes), :330-410 (``aggdg_cg``).  Restriction is always L' (solvers.jl:36).
Every L is element-local: one dense block per (fine element, parent element) pair.
"""
import numpy as np
import scipy.sparse as sp

from .reference_element import evaluate_nodal_basis_fun, gauss_quad


def _blocks_csc(row_nodes, col_nodes, vals, shape, sum_duplicates=True):
    """row_nodes (k, mi), col_nodes (k, mj), vals (k, mi, mj) -> CSC."""
    mi, mj = row_nodes.shape[1], col_nodes.shape[1]
    rows = np.repeat(row_nodes[:, :, None], mj, axis=2).ravel()
    cols = np.repeat(col_nodes[:, None, :], mi, axis=1).ravel()
    return sp.csc_matrix((np.asarray(vals, dtype=np.float64).ravel(), (rows, cols)), shape=shape)


def cg_cg_interpolation(lowMesh, highMesh):
    lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
    hn, ln = highMesh.mNodesInd, lowMesh.mNodesInd
    n = hn.shape[0]
    shape = (highMesh.mNumNodes, lowMesh.mNumNodes)
    parts = []
    if hn.shape[1] > 2:   # interior high nodes: full row of low basis values
        parts.append(_blocks_csc(hn[:, 2:], ln, np.broadcast_to(lowVal[2:, :], (n,) + lowVal[2:, :].shape),
                                 shape))
    # vertices: L[v, v] = lowBasisFunVal[j, j]; assigned (not summed) in the reference (:45-51)
    vr = np.concatenate([hn[:, 0], hn[-1:, 1]])
    vc = np.concatenate([ln[:, 0], ln[-1:, 1]])
    vv = np.concatenate([np.full(n, lowVal[0, 0]), [lowVal[1, 1]]])
    parts.append(sp.csc_matrix((vv, (vr, vc)), shape=shape))
    return sum(parts[1:], parts[0]).tocsc()


def dg_dg_interpolation(lowMesh, highMesh):
    lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
    n = highMesh.mNodesInd.shape[0]
    return _blocks_csc(highMesh.mNodesInd, lowMesh.mNodesInd,
                       np.broadcast_to(lowVal, (n,) + lowVal.shape),
                       (highMesh.mNumNodes, lowMesh.mNumNodes))


def _lumped_solve(N, highMesh):
    """``Diagonal(rowsum(M_CG)) \\ N`` (src/interpolation.jl:210-217)."""
    lumped = np.asarray(highMesh.mMassMatrix.sum(axis=1)).ravel()
    Nc = N.tocoo()
    return sp.csc_matrix((Nc.data / lumped[Nc.row], (Nc.row, Nc.col)), shape=N.shape)


def _half_weights(mesh, nodes):
    """interpFlag 2: 1/2 at interior mesh vertices, 1 at boundary vertices and interior nodes."""
    w = np.ones(nodes.shape)
    for c in (0, 1):
        v = nodes[:, c]
        w[:, c] = np.where((v == 0) | (v == mesh.nVertices - 1), 1.0, 0.5)
    return w


def dg_cg_interpolation(lowMesh, highMesh, mesh, interpFlag):
    hn, ln = highMesh.mNodesInd, lowMesh.mNodesInd
    shape = (highMesh.mNumNodes, lowMesh.mNumNodes)
    if interpFlag in (0, 1):
        gq, gqw = gauss_quad(lowMesh.mP + highMesh.mP)
        highGQ = evaluate_nodal_basis_fun(highMesh.mRefEl.mBasisFunCoeff, gq)
        lowGQ = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, gq)
        ref = np.einsum("l,li,lj->ij", gqw, highGQ, lowGQ)
        N = _blocks_csc(hn, ln, lowMesh.mJacobian[:, None, None] * ref[None], shape)
        if interpFlag == 0:
            return highMesh.mMassMatrixLU.solve(N.toarray())
        return _lumped_solve(N, highMesh)
    if interpFlag == 2:
        lowVal = evaluate_nodal_basis_fun(lowMesh.mRefEl.mBasisFunCoeff, highMesh.mRefEl.mNodesX)
        w = _half_weights(mesh, hn)
        return _blocks_csc(hn, ln, w[:, :, None] * lowVal[None], shape)
    raise ValueError("Only implemented for interpFlag = 0, 1, or 2.")


def aggdg_aggdg_interpolation(coarseMesh, fineMesh, baseMesh):
    if coarseMesh.mP != fineMesh.mP:
        raise ValueError("The two agglomerated meshes must have the same p.")
    n_base = fineMesh.mBaseStarts[-1]
    w = fineMesh.mGaussQuadWeights
    per_base = baseMesh.mJacobian[:n_base, None, None] * np.einsum(
        "l,bli,blj->bij", w, fineMesh.mBasisGQFunVal, coarseMesh.mBasisGQFunVal)
    Nblk = np.add.reduceat(per_base, fineMesh.mBaseStarts[:-1], axis=0)       # one block per fine element
    coarse_of_fine = np.repeat(np.arange(len(coarseMesh.mSubStarts) - 1), np.diff(coarseMesh.mSubStarts))
    N = _blocks_csc(fineMesh.mNodesInd, coarseMesh.mNodesInd[coarse_of_fine], Nblk,
                    (fineMesh.mNumNodes, coarseMesh.mNumNodes))
    return fineMesh.mMassMatrixLU.solve(N)


def aggdg_dg_interpolation(aggMesh, baseMesh):
    n_base = aggMesh.mBaseStarts[-1]
    K = aggMesh.mAggOfBase
    val = aggMesh._basis(baseMesh.mNodesX[:n_base], K[:, None])              # (n_base, m_b, m_agg)
    return _blocks_csc(baseMesh.mNodesInd[:n_base], aggMesh.mNodesInd[K], val,
                       (baseMesh.mNumNodes, aggMesh.mNumNodes))


def aggdg_cg_interpolation(aggMesh, baseMesh, mesh, interpFlag):
    n_base = aggMesh.mBaseStarts[-1]
    K = aggMesh.mAggOfBase
    refEl = baseMesh.mRefEl
    shape = (baseMesh.mNumNodes, aggMesh.mNumNodes)
    if interpFlag in (0, 1):
        gq, gqw = refEl.mGaussQuadNodes, refEl.mGaussQuadWeights
        xq = baseMesh.mXc[:n_base, None] + (baseMesh.mH[:n_base] / 2.0)[:, None] * gq[None, :]
        aggGQ = aggMesh._basis(xq, K[:, None])                                # (n_base, nq, m_agg)
        blk = baseMesh.mJacobian[:n_base, None, None] * np.einsum(
            "l,li,blj->bij", gqw, refEl.mBasisGQFunVal, aggGQ)
        N = _blocks_csc(baseMesh.mNodesInd[:n_base], aggMesh.mNodesInd[K], blk, shape)
        if interpFlag == 0:
            return baseMesh.mMassMatrixLU.solve(N.toarray())
        return _lumped_solve(N, baseMesh)
    if interpFlag == 2:
        val = aggMesh._basis(baseMesh.mNodesX[:n_base], K[:, None])
        w = _half_weights(mesh, baseMesh.mNodesInd[:n_base])
        return _blocks_csc(baseMesh.mNodesInd[:n_base], aggMesh.mNodesInd[K], w[:, :, None] * val, shape)
    raise ValueError("Only implemented for interpFlag = 0, 1, or 2.")
