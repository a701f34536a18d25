"""MeshHierarchy: host mirror of src/mesh_heirarchy.jl (the reference's spelling), plus the upload.

Both reference constructors are kept:
  * ``MeshHierarchy(mMeshes, mesh, mBdConds, A, nCG=, nDG=, nAgg=, CDir=)``   (:30-138, CG first)
  * ``MeshHierarchy(mMeshes, mBdConds, A, G, D, C, nDG=, nAgg=)``             (:140-181, DG first)
The second one accepts ``nAgg`` in the reference but never builds the agglomerated levels; here it
does, mirroring :89-106 (needed by the BASELINE configs; an extension, flagged in DESIGN.md).

Set-up (Galerkin products L'XL, A = C - D (M \\ G), smoother extraction) runs on the host with scipy
exactly as in the reference; at the end of construction every level is converted to element blocks
and uploaded once (``H.device``).  The solver entry points in solvers.py then run on the GPU.
"""
import scipy.sparse as sp

from . import blocks as blk
from .agglomerated_dg_mesh import AgglomeratedDgMesh1, AgglomeratedDgMeshN, agg_dg_flux_operators
from .cg_mesh import CgMesh
from .device import DeviceHierarchy
from .dg_mesh import DgMesh
from .dg_mesh import dg_flux_operators as _dg_flux_operators
from .interpolation import (aggdg_aggdg_interpolation, aggdg_cg_interpolation,
                            aggdg_dg_interpolation, cg_cg_interpolation, dg_cg_interpolation,
                            dg_dg_interpolation)
from .smoother import cg_smoother, dg_smoother, smoother_inverse


def dg_flux_operators(dgMesh, mesh_or_base, bdCond, CDir):
    """Dispatch like the reference's two methods (src/dg_mesh.jl:144, agglomerated_dg_mesh.jl:641)."""
    if isinstance(dgMesh, AgglomeratedDgMesh1):
        return agg_dg_flux_operators(dgMesh, mesh_or_base, bdCond, CDir)
    if isinstance(dgMesh, DgMesh):
        return _dg_flux_operators(dgMesh, mesh_or_base, bdCond, CDir)
    raise TypeError("dg_flux_operators needs a DgMesh or an AgglomeratedDgMesh1")


def _galerkin(L, X):
    return (L.T @ X @ L).tocsc()


def _dg_level(mesh, G, D, C):
    A = (C - D @ mesh.mMassMatrixLU.solve(G)).tocsc()
    return A, dg_smoother(mesh, A, "blockJac")


class MeshHierarchy:
    """Fields as in src/mesh_heirarchy.jl:17-28: mMeshes, mStiffness, mGradient, mDivergence, mC,
    mSmoothers, mInterpolation, mBdConds; plus ``device`` (the uploaded hierarchy)."""

    def __init__(self, mMeshes, *args, nCG=None, nDG=None, nAgg=0, CDir=1.0, device=0, upload=True,
                 stream=None, device_setup=False):
        """device_setup=True: the host assembles what the reference assembles from the mesh (level 0's
        operator, the flux operators of the first DG-type level, the element-local transfers); every
        Galerkin product - L' A L of the CG levels, L' (G, D, C) L of the DG-type levels -, A = C - D (M \\ G)
        and the smoothers of every level are computed on the GPU.  mStiffness[k], k > 0, then
        stay None on the host; ``level_blocks(k)`` downloads them."""
        self.mMeshes = list(mMeshes)
        self.device_setup = bool(device_setup)
        if len(args) == 3:                      # (mesh, mBdConds, A): CG-first constructor
            mesh, mBdConds, A = args
            if self.device_setup:
                self._build_cg_first_on_device(mesh, mBdConds, A, 1 if nCG is None else nCG,
                                               0 if nDG is None else nDG, nAgg, CDir, device, stream)
                return
            self._build_cg_first(mesh, mBdConds, A, 1 if nCG is None else nCG,
                                 0 if nDG is None else nDG, nAgg, CDir)
        elif len(args) == 5:                    # (mBdConds, A, G, D, C): DG-first constructor
            mBdConds, A, G, D, C = args
            if self.device_setup:
                self._build_dg_first_on_device(mBdConds, A, G, D, C, 1 if nDG is None else nDG, nAgg,
                                               device, stream)
                return
            self._build_dg_first(mBdConds, A, G, D, C, 1 if nDG is None else nDG, nAgg)
        else:
            raise TypeError("MeshHierarchy(mMeshes, mesh, mBdConds, A; ...) or "
                            "MeshHierarchy(mMeshes, mBdConds, A, G, D, C; ...)")
        self.device = None
        if upload:
            self.upload(device=device, stream=stream)

    # ---- src/mesh_heirarchy.jl:30-138 -----------------------------------------------------------
    def _build_cg_first(self, mesh, mBdConds, A, nCG, nDG, nAgg, CDir):
        M = self.mMeshes
        if nCG <= 0:
            raise ValueError("At least one CG mesh required.")
        if len(M) != nCG + nDG + nAgg:
            raise ValueError("Length of vector of meshes does not match inputed number of CG, DG, "
                             "and agglomerated meshes.")
        nL = nCG + nDG + nAgg
        S, Sm, I = [None] * nL, [None] * nL, [None] * (nL - 1)
        Gs, Ds, Cs = ([None] * (nDG + nAgg) for _ in range(3))
        S[0] = sp.csc_matrix(A)
        Sm[0] = cg_smoother(M[0], S[0], "jac")
        for i in range(1, nCG):
            I[i - 1] = cg_cg_interpolation(M[i], M[i - 1])
            S[i] = _galerkin(I[i - 1], S[i - 1])
            Sm[i] = cg_smoother(M[i], S[i], "jac")

        def project(g, L):
            Gs[g], Ds[g], Cs[g] = (_galerkin(L, Gs[g - 1]), _galerkin(L, Ds[g - 1]),
                                   _galerkin(L, Cs[g - 1]))

        if nDG >= 1:
            I[nCG - 1] = dg_cg_interpolation(M[nCG], M[nCG - 1], mesh, 1)
            Gs[0], Ds[0], Cs[0] = dg_flux_operators(M[nCG], mesh, mBdConds[nCG], CDir)
            S[nCG], Sm[nCG] = _dg_level(M[nCG], Gs[0], Ds[0], Cs[0])
            for i in range(1, nDG):
                I[nCG + i - 1] = dg_dg_interpolation(M[nCG + i], M[nCG + i - 1])
                project(i, I[nCG + i - 1])
                S[nCG + i], Sm[nCG + i] = _dg_level(M[nCG + i], Gs[i], Ds[i], Cs[i])
            for i in range(nAgg):
                k = nCG + nDG + i
                if i == 0:
                    I[k - 1] = aggdg_dg_interpolation(M[k], M[k - 1])
                else:
                    I[k - 1] = aggdg_aggdg_interpolation(M[k], M[k - 1], M[nCG + nDG - 1])
                project(nDG + i, I[k - 1])
                S[k], Sm[k] = _dg_level(M[k], Gs[nDG + i], Ds[nDG + i], Cs[nDG + i])
        elif nAgg >= 1:
            I[nCG - 1] = aggdg_cg_interpolation(M[nCG], M[nCG - 1], mesh, 1)
            Gs[0], Ds[0], Cs[0] = dg_flux_operators(M[nCG], M[nCG - 1], mBdConds[nCG], CDir)
            S[nCG], Sm[nCG] = _dg_level(M[nCG], Gs[0], Ds[0], Cs[0])
            for i in range(1, nAgg):
                I[nCG + i - 1] = aggdg_aggdg_interpolation(M[nCG + i], M[nCG + i - 1], M[nCG - 1])
                project(i, I[nCG + i - 1])
                S[nCG + i], Sm[nCG + i] = _dg_level(M[nCG + i], Gs[i], Ds[i], Cs[i])
        self._set(S, Gs, Ds, Cs, Sm, I, mBdConds)

    # ---- src/mesh_heirarchy.jl:140-181 (+ agglomerated tail) -------------------------------------
    def _build_dg_first(self, mBdConds, A, G, D, C, nDG, nAgg):
        M = self.mMeshes
        if nDG <= 0:
            raise ValueError("At least one DG mesh required.")
        if len(M) != nDG + nAgg:
            raise ValueError("Length of vector of meshes does not match inputed number of DG and "
                             "agglomerated meshes.")
        nL = nDG + nAgg
        S, Sm, I = [None] * nL, [None] * nL, [None] * (nL - 1)
        Gs, Ds, Cs = ([None] * nL for _ in range(3))
        Gs[0], Ds[0], Cs[0] = sp.csc_matrix(G), sp.csc_matrix(D), sp.csc_matrix(C)
        S[0] = sp.csc_matrix(A)
        Sm[0] = dg_smoother(M[0], S[0], "blockJac")
        for k in range(1, nL):
            if k < nDG:
                I[k - 1] = dg_dg_interpolation(M[k], M[k - 1])
            elif k == nDG:
                I[k - 1] = aggdg_dg_interpolation(M[k], M[k - 1])
            else:
                I[k - 1] = aggdg_aggdg_interpolation(M[k], M[k - 1], M[nDG - 1])
            L = I[k - 1]
            Gs[k], Ds[k], Cs[k] = _galerkin(L, Gs[k - 1]), _galerkin(L, Ds[k - 1]), _galerkin(L, Cs[k - 1])
            S[k], Sm[k] = _dg_level(M[k], Gs[k], Ds[k], Cs[k])
        self._set(S, Gs, Ds, Cs, Sm, I, mBdConds)

    # ---- the same chain with the Galerkin products on the GPU ------------------------
    def _build_dg_first_on_device(self, mBdConds, A, G, D, C, nDG, nAgg, device, stream):
        from .smoother import DeviceBlockJacobi
        import numpy as np
        M = self.mMeshes
        if nDG <= 0:
            raise ValueError("At least one DG mesh required.")
        if len(M) != nDG + nAgg:
            raise ValueError("Length of vector of meshes does not match inputed number of DG and "
                             "agglomerated meshes.")
        nL = nDG + nAgg
        slots = [blk.level_slots(m) for m in M]
        if not all(blk.is_identity_slots(s) for s in slots):
            raise ValueError("device-side set-up needs DG-type levels (element-major DOF numbering)")
        dev = DeviceHierarchy(nL, device=device, stream=stream)
        minv = lambda mesh: np.linalg.inv(mesh.mMassMatrix.mBlocks)          # noqa: E731
        flux = [blk.csc_to_blocks(sp.csc_matrix(X), slots[0], pad_identity=False) for X in (G, D, C)]
        dev.set_level_flux(0, flux[0], flux[1], flux[2], minv(M[0]))
        I = [None] * (nL - 1)
        for k in range(1, nL):
            if k < nDG:
                I[k - 1] = dg_dg_interpolation(M[k], M[k - 1])
            elif k == nDG:
                I[k - 1] = aggdg_dg_interpolation(M[k], M[k - 1])
            else:
                I[k - 1] = aggdg_aggdg_interpolation(M[k], M[k - 1], M[nDG - 1])
            parent, P0, P1 = blk.transfer_to_blocks(I[k - 1], slots[k - 1], slots[k])
            dev.set_transfer_blocks(k - 1, parent, P0, P1)
            dev.coarsen_level(k - 1, minv(M[k]), slots[k].shape[0])
        dev.finalize()
        S = [sp.csc_matrix(A)] + [None] * (nL - 1)
        Sm = [DeviceBlockJacobi(dev, l) for l in range(nL)]
        none = [None] * nL
        self._set(S, [sp.csc_matrix(G)] + none[1:], [sp.csc_matrix(D)] + none[1:],
                  [sp.csc_matrix(C)] + none[1:], Sm, I, mBdConds)
        self.device = dev
        self.mSlots = slots

    # ---- the CG-first chain with every product on the GPU ------------------------------
    def _build_cg_first_on_device(self, mesh, mBdConds, A, nCG, nDG, nAgg, CDir, device, stream):
        """src/mesh_heirarchy.jl:30-138 with mStiffness[i] = L' mStiffness[i-1] L of the CG levels
        (amg1d_coarsen_level_galerkin, two-parent cg_cg transfers) and the DG-type chain underneath
        (amg1d_set_level_flux on its first level, amg1d_coarsen_level below) formed on the device."""
        from .smoother import DeviceBlockJacobi, DeviceJacobi
        import numpy as np
        M = self.mMeshes
        if nCG <= 0:
            raise ValueError("At least one CG mesh required.")
        if len(M) != nCG + nDG + nAgg:
            raise ValueError("Length of vector of meshes does not match inputed number of CG, DG, "
                             "and agglomerated meshes.")
        nL = nCG + nDG + nAgg
        slots = [blk.level_slots(m) for m in M]
        dev = DeviceHierarchy(nL, device=device, stream=stream)
        S0 = sp.csc_matrix(A)
        Sm = [None] * nL
        Sm[0] = cg_smoother(M[0], S0, "jac")
        lo, di, up = blk.csc_to_blocks(S0, slots[0])
        dinv, is_diag = smoother_inverse(Sm[0], slots[0])
        dev.set_level_blocks(0, lo, di, up, dinv, is_diag, slots[0], S0.shape[0])
        Sm[0]._owner = (dev, 0)
        I = [None] * (nL - 1)

        def transfer(k, L):
            I[k] = L
            parent, P0, P1 = blk.transfer_to_blocks(L, slots[k], slots[k + 1])
            dev.set_transfer_blocks(k, parent, P0, P1)

        for i in range(1, nCG):
            transfer(i - 1, cg_cg_interpolation(M[i], M[i - 1]))
            dev.coarsen_level_galerkin(i - 1, slots[i], M[i].mNumNodes, dinv_is_diagonal=True)
            Sm[i] = DeviceJacobi(dev, i, slots[i], M[i].mNumNodes)
        minv = lambda m: np.linalg.inv(m.mMassMatrix.mBlocks)                # noqa: E731
        G = D = C = None
        if nDG + nAgg >= 1:
            if nDG >= 1:
                transfer(nCG - 1, dg_cg_interpolation(M[nCG], M[nCG - 1], mesh, 1))
                G, D, C = dg_flux_operators(M[nCG], mesh, mBdConds[nCG], CDir)
            else:
                transfer(nCG - 1, aggdg_cg_interpolation(M[nCG], M[nCG - 1], mesh, 1))
                G, D, C = dg_flux_operators(M[nCG], M[nCG - 1], mBdConds[nCG], CDir)
            flux = [blk.csc_to_blocks(sp.csc_matrix(X), slots[nCG], pad_identity=False) for X in (G, D, C)]
            dev.set_level_flux(nCG, flux[0], flux[1], flux[2], minv(M[nCG]))
            for k in range(nCG + 1, nL):
                if k < nCG + nDG:
                    L = dg_dg_interpolation(M[k], M[k - 1])
                elif k == nCG + nDG:
                    L = aggdg_dg_interpolation(M[k], M[k - 1])
                else:
                    L = aggdg_aggdg_interpolation(M[k], M[k - 1], M[nCG + nDG - 1] if nDG else M[nCG - 1])
                transfer(k - 1, L)
                dev.coarsen_level(k - 1, minv(M[k]), slots[k].shape[0])
            for k in range(nCG, nL):
                Sm[k] = DeviceBlockJacobi(dev, k)
        dev.finalize()
        nD = nDG + nAgg
        none = [None] * max(nD, 1)
        first = lambda X: ([sp.csc_matrix(X)] + none[1:nD]) if nD else []    # noqa: E731
        self._set([S0] + [None] * (nL - 1), first(G), first(D), first(C), Sm, I, mBdConds)
        self.device = dev
        self.mSlots = slots

    def level_blocks(self, l):
        """(lo, di, up, dinv) of level l as the device holds them (downloads; any set-up path)."""
        n, m = self.mSlots[l].shape
        from .smoother import JacobiSmoother
        return self.device.get_level(l, n, m, diag=isinstance(self.mSmoothers[l], JacobiSmoother))

    def _set(self, S, Gs, Ds, Cs, Sm, I, mBdConds):
        self.mStiffness, self.mGradient, self.mDivergence, self.mC = S, Gs, Ds, Cs
        self.mSmoothers, self.mInterpolation, self.mBdConds = Sm, I, list(mBdConds)

    # ---- upload (stands in for the end of construction) ------------------------------------------
    def upload(self, device=0, stream=None, options=None, dist=None):
        """options: dict for amg1d_set_option, applied before the first level (e.g. {"compress": 0}).
        dist = (rank, nranks, nccl_id_bytes): every rank passes the same global arrays and the library keeps its
        contiguous slab of the large levels (DG-first hierarchies; any mesh - the transfers go up as explicit
        per-element blocks); host vectors are then the rank's slab (``device.info("local_dofs")``)."""
        nL = len(self.mMeshes)
        dev = DeviceHierarchy(nL, device=device, stream=stream, dist=dist)
        for k, v in (options or {}).items():
            dev.set_option(k, v)
        slots = [blk.level_slots(m) for m in self.mMeshes]
        for l in range(nL):
            lo, di, up = blk.csc_to_blocks(self.mStiffness[l], slots[l])
            dinv, is_diag = smoother_inverse(self.mSmoothers[l], slots[l])
            dev.set_level_blocks(l, lo, di, up, dinv, is_diag, slots[l], self.mStiffness[l].shape[0])
            self.mSmoothers[l]._owner = (dev, l)
        for l in range(nL - 1):
            parent, P0, P1 = blk.transfer_to_blocks(self.mInterpolation[l], slots[l], slots[l + 1])
            dev.set_transfer_blocks(l, parent, P0, P1)
        dev.finalize()
        if dist is not None and dist[1] > 1:
            dev.n_dof[0] = dev.info("local_dofs")
        self.device = dev
        self.mSlots = slots
        return dev
