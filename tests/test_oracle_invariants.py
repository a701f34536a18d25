"""Pins for the CPU oracle.  The reference stores no golden vectors (its tests println and plot),
so the oracle is pinned by the invariants the reference scripts print:
Galerkin consistency ~ 0, exact reproduction of coarse-space polynomials, discretisation order
~ p+1, operator structure, mesh-independent V-cycle counts."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import aggdg, cg, dg, drivers, interpolation as interp, refelem, refmesh, solvers


def _mesh(n, bc=("neu", "dir"), xin=0.0, xout=1.0):
    mesh = refmesh.create_uniform_mesh(n, xin, xout)
    vals = [(k, -math.sin(x) if k == "neu" else math.cos(x)) for k, x in zip(bc, (xin, xout))]
    return mesh, refmesh.set_boundary(mesh, xin, xout, vals)


def test_reference_element_and_quadrature():
    for p in range(0, 9):
        re = refelem.ReferenceElement(p)
        # nodal basis: phi_j(x_i) = delta_ij
        V = refelem.evaluate_nodal_basis_fun(re.mBasisFunCoeff, re.mNodesX)
        assert np.allclose(V, np.eye(p + 1), atol=1e-12)
        # Gauss rule of degree 2p integrates x^k exactly for k <= 2p
        x, w = re.mGaussQuadNodes, re.mGaussQuadWeights
        assert len(x) == p + 1
        for k in range(2 * p + 1):
            exact = 0.0 if k % 2 else 2.0 / (k + 1)
            assert abs(np.dot(w, x ** k) - exact) < 1e-13
        assert abs(re.mMassMatrix.sum() - 2.0) < 1e-13      # partition of unity


def test_dg_operator_structure():
    """F8: A symmetric positive definite, block tridiagonal, G = -D' away from the boundary terms."""
    for p in (1, 3):
        n = 8
        mesh, bd = _mesh(n)
        m = dg.DgMesh(mesh, p)
        A, b, G, D, C = dg.dg_operator_and_rhs(m, mesh, math.cos, bd, 1000.0 * n)
        Ad = A.toarray()
        assert np.abs(Ad - Ad.T).max() < 1e-9 * np.abs(Ad).max()
        assert np.linalg.eigvalsh((Ad + Ad.T) / 2).min() > 0
        k = p + 1
        for e in range(n):
            for f in range(n):
                if abs(e - f) > 1:
                    assert np.all(Ad[e * k:(e + 1) * k, f * k:(f + 1) * k] == 0.0)
        # interior blocks identical on a uniform mesh
        assert np.allclose(Ad[2 * k:3 * k, 2 * k:3 * k], Ad[4 * k:5 * k, 4 * k:5 * k], rtol=1e-12)
        # G + D' vanishes except for the Neumann (G) and Dirichlet (D) boundary entries
        R = (G + D.T).toarray()
        R[0, 0] = 0.0
        R[(n - 1) * k + 1, (n - 1) * k + 1] = 0.0
        assert np.abs(R).max() < 1e-13


def test_galerkin_consistency_dg():
    """tests/dg_interpolation_test.jl:40-44: || lowX - L' highX L || ~ 0 for X in {G, D, C, M}."""
    n = 16
    mesh, bd = _mesh(n, bc=("dir", "dir"))
    high, low = dg.DgMesh(mesh, 4), dg.DgMesh(mesh, 2)
    L = interp.dg_dg_interpolation(low, high)
    Gh, Dh, Ch = dg.dg_flux_operators(high, mesh, bd, 1000.0 * n)
    Gl, Dl, Cl = dg.dg_flux_operators(low, mesh, bd, 1000.0 * n)
    for Xl, Xh in ((Gl, Gh), (Dl, Dh), (Cl, Ch), (low.mMassMatrix.tosparse(), high.mMassMatrix.tosparse())):
        assert abs(Xl - L.T @ Xh @ L).max() < 1e-11 * max(1.0, abs(Xh).max())


def test_galerkin_consistency_aggdg_dg():
    """tests/aggdg_dg_interpolation_test.jl:46-50 with a DG p=1 base and pAgg = 1, factor 2."""
    n = 16
    mesh, bd = _mesh(n, bc=("dir", "neu"))
    base = dg.DgMesh(mesh, 1)
    amap = drivers.agglomeration_maps(n, [2])[0]
    am = aggdg.AgglomeratedDgMesh1(1, amap, mesh, base)
    L = interp.aggdg_dg_interpolation(am, base)
    Gh, Dh, Ch = dg.dg_flux_operators(base, mesh, bd, 100.0)
    Gl, Dl, Cl = aggdg.agg_dg_flux_operators(am, base, bd, 100.0)
    for Xl, Xh in ((Gl, Gh), (Dl, Dh), (Cl, Ch), (am.mMassMatrix.tosparse(), base.mMassMatrix.tosparse())):
        assert abs(Xl - L.T @ Xh @ L).max() < 1e-11 * max(1.0, abs(Xh).max())


def test_galerkin_consistency_aggdg_aggdg():
    """tests/aggdg_interpolation_test.jl:59-63: mass matrix of the coarser agglomerated mesh equals
    the projection of the finer one."""
    n = 16
    mesh, bd = _mesh(n)
    base = dg.DgMesh(mesh, 1)
    maps = drivers.agglomeration_maps(n, [2, 2])
    fine = aggdg.AgglomeratedDgMesh1(1, maps[0], mesh, base)
    coarse = aggdg.AgglomeratedDgMeshN(1, maps[1], fine, base)
    L = interp.aggdg_aggdg_interpolation(coarse, fine, base)
    Mf, Mc = fine.mMassMatrix.tosparse(), coarse.mMassMatrix.tosparse()
    assert abs(Mc - L.T @ Mf @ L).max() < 1e-13
    # element maps: mBaseElementInds is the concatenation of the sub-elements' lists (:507-511)
    assert coarse.mElements[1].mBaseElementInds == [5, 6, 7, 8]
    assert coarse.mElements[1].mSubAggElementInds == [3, 4]
    assert list(coarse.mElements[1].mNodesInd) == [2, 3]


def test_galerkin_consistency_cg():
    """tests/cg_interpolation_test.jl:43: lowA = L' highA L (before the Dirichlet modification)."""
    n = 8
    mesh, bd = _mesh(n, bc=("neu", "neu"))
    high, low = cg.CgMesh(mesh, 4), cg.CgMesh(mesh, 2)
    L = interp.cg_cg_interpolation(low, high)
    Ah, Al = cg.cg_stiffness(high, bd), cg.cg_stiffness(low, bd)
    assert abs(Al - L.T @ Ah @ L).max() < 1e-11 * abs(Ah).max()


def test_prolongation_reproduces_coarse_polynomials():
    n = 8
    mesh, _ = _mesh(n)
    high, low = dg.DgMesh(mesh, 4), dg.DgMesh(mesh, 2)
    L = interp.dg_dg_interpolation(low, high)
    f = lambda x: 1.0 + 2.0 * x - 3.0 * x * x
    ul = np.concatenate([f(el.mNodesX) for el in low.mElements])
    uh = np.concatenate([f(el.mNodesX) for el in high.mElements])
    assert np.abs(L @ ul - uh).max() < 1e-12
    # CG -> DG lumped L2 projection reproduces constants
    cgm, dgm = cg.CgMesh(mesh, 2), dg.DgMesh(mesh, 1)
    Lc = interp.dg_cg_interpolation(dgm, cgm, mesh, 1)
    assert np.abs(Lc @ np.ones(dgm.mNumNodes) - 1.0).max() < 1e-12


@pytest.mark.parametrize("kind,p", [("dg", 1), ("dg", 2), ("cg", 1), ("cg", 2)])
def test_discretisation_order(kind, p):
    """tests/cg_convergence_test.jl / dg_convergence_test.jl: error ~ h^(p+1)."""
    errs = []
    ns = [8, 16, 32]
    for n in ns:
        mesh, bd = _mesh(n, bc=("dir", "neu"))
        if kind == "dg":
            m = dg.DgMesh(mesh, p)
            A, b, *_ = dg.dg_operator_and_rhs(m, mesh, math.cos, bd, 10.0 * n)
        else:
            m = cg.CgMesh(mesh, p)
            A, b = cg.cg_stiffness_and_rhs(m, mesh, math.cos, bd)
        u = spla.spsolve(sp.csc_matrix(A), b)
        e = 0.0
        for el in m.mElements:
            e = max(e, np.abs(u[el.mNodesInd] - np.cos(el.mNodesX)).max())
        errs.append(e)
    rates = [math.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    assert min(rates) > p + 0.5, (errs, rates)


def test_aggdg_discretisation_converges():
    """tests/aggdg_convergence_test.jl shape: DG p=1 base, pAgg=0, factor 2, Dirichlet-left /
    Neumann-right, CDir = n: first-order convergence of the piecewise-constant solution."""
    errs = []
    for n in (16, 32, 64):
        mesh, bd = _mesh(n, bc=("dir", "neu"))
        base = dg.DgMesh(mesh, 1)
        am = aggdg.AgglomeratedDgMesh1(0, drivers.agglomeration_maps(n, [2])[0], mesh, base)
        G, D, C = aggdg.agg_dg_flux_operators(am, base, bd, 1.0 * n)
        f, r = aggdg.agg_dg_flux_rhs(am, base, math.cos, bd, 1.0 * n)
        A = (C - D @ am.mMassMatrixLU.solve(G)).tocsc()
        b = f - D @ am.mMassMatrixLU.solve(r)
        u = spla.spsolve(A, b)
        xc = np.array([0.5 * (el.mBoundingBox[0] + el.mBoundingBox[1]) for el in am.mElements])
        errs.append(np.abs(u - np.cos(xc)).max())
    assert errs[1] < 0.7 * errs[0] and errs[2] < 0.7 * errs[1], errs


def test_vcycle_counts_mesh_independent():
    """tests/full_heirarchy_test.jl:98 prints iter for n = 8..512; the count must not grow with n.
    Also the DG -> agglomerated shape of the BASELINE configs (11 cycles)."""
    its = []
    for n in (8, 16, 32, 64):
        H, x0, b, _ = drivers.full_heirarchy_test(n=n)
        its.append(solvers.multigrid(H, x0, b, 100, 1e-10)[1])
    assert max(its) - min(its) <= 2 and max(its) <= 14, its
    its2 = []
    for n in (32, 128):
        H, x0, b, _ = drivers.dg_agg_problem(n, p=3)
        its2.append(solvers.multigrid(H, x0, b, 100, 1e-10)[1])
    assert its2[0] == its2[1] == 11, its2
    H, x0, b, _ = drivers.dg_heirarchy_test(n=32)
    assert solvers.multigrid(H, x0, b, 200, 1e-10)[1] == 9


def test_batched_solve_equals_lu_loop():
    """apply_smoother's per-element getrs loop == one batched LAPACK gesv per block."""
    H, x0, b, _ = drivers.dg_heirarchy_test(n=8)
    from oracle import smoother as osm
    S = H.mSmoothers[0]
    y = osm.apply_smoother(S, b, alpha=1.0)
    A = sp.csc_matrix(H.mStiffness[0])
    idx = S.mBlockInds.T
    blocks = np.stack([A[i, :][:, i].toarray() for i in idx])
    y2 = np.zeros_like(b)
    y2[idx] = np.linalg.solve(blocks, b[idx][:, :, None])[:, :, 0]
    assert np.abs(y - y2).max() <= 1e-14 * np.abs(y).max()


def test_c_restatement_matches_python_oracle():
    """oracle/vcycle_ref.c (the timed CPU baseline) against the literal Python oracle."""
    from oracle import cref
    for build in (lambda: drivers.dg_agg_problem(64, p=3, unit_h=True), lambda: drivers.full_heirarchy_test(n=32),
                  lambda: drivers.dg_cg_heirarchy_test(n=16)):
        H, x0, b, _ = build()
        c = cref.CRefHierarchy(H)
        rng = np.random.default_rng(0)
        xs = rng.standard_normal(len(b))
        for nPre, nPost, alpha in ((3, 3, 2.0 / 3.0), (1, 2, 0.5)):
            x_py = solvers.multigrid_v_cycle(H, xs, b, nPre=nPre, nPost=nPost, alpha=alpha)
            x_c = c.vcycle(xs, b, nPre=nPre, nPost=nPost, alpha=alpha)
            assert np.abs(x_c - x_py).max() <= 1e-11 * np.abs(x_py).max()
        assert abs(c.residual_norm(x_c, b) - np.linalg.norm(H.mStiffness[0] @ x_c - b)) <= 1e-9 * np.linalg.norm(b)
        c.close()
