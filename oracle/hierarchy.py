"""Oracle (test infrastructure): the multigrid hierarchy (data contract of the V-cycle).

Follows src/mesh_heirarchy.jl:17-28 (struct), :30-138 (CG-first constructor), :140-181 (DG-first
constructor).  The DG-first constructor of the reference accepts ``nAgg`` but never builds the
agglomerated levels; ``MeshHierarchyDG`` here adds that loop as a mirror of :89-106 (this is an
extension the BASELINE configs need, marked as such in DESIGN.md).
"""
import scipy.sparse as sp

from .aggdg import agg_dg_flux_operators
from .dg import dg_flux_operators
from .interpolation import (aggdg_aggdg_interpolation, aggdg_cg_interpolation,
                            aggdg_dg_interpolation, cg_cg_interpolation, dg_cg_interpolation,
                            dg_dg_interpolation)
from .smoother import cg_smoother, dg_smoother


class MeshHierarchy:
    def __init__(self, mMeshes, mStiffness, mGradient, mDivergence, mC, mSmoothers,
                 mInterpolation, mBdConds):
        self.mMeshes = mMeshes
        self.mStiffness = mStiffness
        self.mGradient = mGradient
        self.mDivergence = mDivergence
        self.mC = mC
        self.mSmoothers = mSmoothers
        self.mInterpolation = mInterpolation
        self.mBdConds = mBdConds


def _galerkin(L, X):
    return (L.T @ X @ L).tocsc()


def _dg_level(mesh, G, D, C):
    """A = C - D (M_LU \\ G) with the level's own mass matrix, plus its block-Jacobi smoother."""
    A = (C - D @ mesh.mMassMatrixLU.solve(G)).tocsc()
    return A, dg_smoother(mesh, A, "blockJac")


def MeshHierarchyCG(mMeshes, mesh, mBdConds, A, nCG=1, nDG=0, nAgg=0, CDir=1.0):
    """``MeshHierarchy(mMeshes, mesh, mBdConds, A; nCG, nDG, nAgg, CDir)`` (:30-138)."""
    if nCG <= 0:
        raise ValueError("At least one CG mesh required.")
    if len(mMeshes) != nCG + nDG + nAgg:
        raise ValueError("Length of vector of meshes does not match inputed number of CG, DG, "
                         "and agglomerated meshes.")
    nL = nCG + nDG + nAgg
    S = [None] * nL
    Gs = [None] * (nDG + nAgg)
    Ds = [None] * (nDG + nAgg)
    Cs = [None] * (nDG + nAgg)
    Sm = [None] * nL
    I = [None] * (nL - 1)
    S[0] = sp.csc_matrix(A)
    Sm[0] = cg_smoother(mMeshes[0], S[0], "jac")
    for i in range(1, nCG):
        L = cg_cg_interpolation(mMeshes[i], mMeshes[i - 1])
        I[i - 1] = L
        S[i] = _galerkin(L, S[i - 1])
        Sm[i] = cg_smoother(mMeshes[i], S[i], "jac")
    if nDG >= 1:
        I[nCG - 1] = dg_cg_interpolation(mMeshes[nCG], mMeshes[nCG - 1], mesh, 1)
        Gs[0], Ds[0], Cs[0] = dg_flux_operators(mMeshes[nCG], mesh, mBdConds[nCG], CDir)
        S[nCG], Sm[nCG] = _dg_level(mMeshes[nCG], Gs[0], Ds[0], Cs[0])
        for i in range(1, nDG):
            L = dg_dg_interpolation(mMeshes[nCG + i], mMeshes[nCG + i - 1])
            I[nCG + i - 1] = L
            Gs[i], Ds[i], Cs[i] = (_galerkin(L, Gs[i - 1]), _galerkin(L, Ds[i - 1]),
                                   _galerkin(L, Cs[i - 1]))
            S[nCG + i], Sm[nCG + i] = _dg_level(mMeshes[nCG + i], Gs[i], Ds[i], Cs[i])
        for i in range(nAgg):
            k = nCG + nDG + i
            if i == 0:
                L = aggdg_dg_interpolation(mMeshes[k], mMeshes[k - 1])
            else:
                L = aggdg_aggdg_interpolation(mMeshes[k], mMeshes[k - 1], mMeshes[nCG + nDG - 1])
            I[k - 1] = L
            g = nDG + i
            Gs[g], Ds[g], Cs[g] = (_galerkin(L, Gs[g - 1]), _galerkin(L, Ds[g - 1]),
                                   _galerkin(L, Cs[g - 1]))
            S[k], Sm[k] = _dg_level(mMeshes[k], Gs[g], Ds[g], Cs[g])
    elif nAgg >= 1:
        I[nCG - 1] = aggdg_cg_interpolation(mMeshes[nCG], mMeshes[nCG - 1], mesh, 1)
        Gs[0], Ds[0], Cs[0] = agg_dg_flux_operators(mMeshes[nCG], mMeshes[nCG - 1],
                                                    mBdConds[nCG], CDir)
        S[nCG], Sm[nCG] = _dg_level(mMeshes[nCG], Gs[0], Ds[0], Cs[0])
        for i in range(1, nAgg):
            L = aggdg_aggdg_interpolation(mMeshes[nCG + i], mMeshes[nCG + i - 1], mMeshes[nCG - 1])
            I[nCG + i - 1] = L
            Gs[i], Ds[i], Cs[i] = (_galerkin(L, Gs[i - 1]), _galerkin(L, Ds[i - 1]),
                                   _galerkin(L, Cs[i - 1]))
            S[nCG + i], Sm[nCG + i] = _dg_level(mMeshes[nCG + i], Gs[i], Ds[i], Cs[i])
    return MeshHierarchy(mMeshes, S, Gs, Ds, Cs, Sm, I, mBdConds)


def MeshHierarchyDG(mMeshes, mBdConds, A, G, D, C, nDG=1, nAgg=0):
    """``MeshHierarchy(mMeshes, mBdConds, A, G, D, C; nDG, nAgg)`` (:140-181) plus the
    agglomerated tail the reference leaves out (mirror of :89-106)."""
    if nDG <= 0:
        raise ValueError("At least one DG mesh required.")
    if len(mMeshes) != nDG + nAgg:
        raise ValueError("Length of vector of meshes does not match inputed number of DG and "
                         "agglomerated meshes.")
    nL = nDG + nAgg
    S = [None] * nL
    Gs = [None] * nL
    Ds = [None] * nL
    Cs = [None] * nL
    Sm = [None] * nL
    I = [None] * (nL - 1)
    Gs[0], Ds[0], Cs[0] = sp.csc_matrix(G), sp.csc_matrix(D), sp.csc_matrix(C)
    S[0] = sp.csc_matrix(A)
    Sm[0] = dg_smoother(mMeshes[0], S[0], "blockJac")
    for i in range(1, nDG):
        L = dg_dg_interpolation(mMeshes[i], mMeshes[i - 1])
        I[i - 1] = L
        Gs[i], Ds[i], Cs[i] = (_galerkin(L, Gs[i - 1]), _galerkin(L, Ds[i - 1]),
                               _galerkin(L, Cs[i - 1]))
        S[i], Sm[i] = _dg_level(mMeshes[i], Gs[i], Ds[i], Cs[i])
    for i in range(nAgg):
        k = nDG + i
        if i == 0:
            L = aggdg_dg_interpolation(mMeshes[k], mMeshes[k - 1])
        else:
            L = aggdg_aggdg_interpolation(mMeshes[k], mMeshes[k - 1], mMeshes[nDG - 1])
        I[k - 1] = L
        Gs[k], Ds[k], Cs[k] = (_galerkin(L, Gs[k - 1]), _galerkin(L, Ds[k - 1]),
                               _galerkin(L, Cs[k - 1]))
        S[k], Sm[k] = _dg_level(mMeshes[k], Gs[k], Ds[k], Cs[k])
    return MeshHierarchy(mMeshes, S, Gs, Ds, Cs, Sm, I, mBdConds)
