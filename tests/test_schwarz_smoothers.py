"""Overlapping Schwarz smoothers of CG levels (src/smoother.jl:1-46, :104-135), the shape of
tests/cg_smoother_test.jl (n = 16, p = 4, Dirichlet at both ends, f = 1).

CPU: in the [vertex_k, interior_k] grouping the smoother is a block-tridiagonal operator; the assembled
operator must equal the oracle's element-loop ``apply_smoother`` applied to the identity.
GPU: apply_smoother (vector, dense and sparse-matrix right-hand sides), ten damped smoothing steps on
sine modes, and iterative_smoother_solve against the oracle."""
import numpy as np
import pytest

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import blocks as blk
from agglomerationmultigrid1d_b200.smoother import schwarz_tridiag_blocks
from oracle import cg as ocg, refmesh, smoother as osm, solvers as osolv


def _problem(n=16, p=4, func=lambda x: 1.0, u=lambda x: -0.5 * x * x + x):
    mesh = aggmg.create_uniform_mesh(n, 0.0, 1.0)
    bd = aggmg.set_boundary(mesh, 0.0, 1.0, [("dir", u(0.0)), ("dir", u(1.0))])
    cgm = aggmg.CgMesh(mesh, p)
    A, b = aggmg.cg_stiffness_and_rhs(cgm, mesh, lambda x: func(x) + 0 * x, bd)
    mo = refmesh.create_uniform_mesh(n, 0.0, 1.0)
    bo = refmesh.set_boundary(mo, 0.0, 1.0, [("dir", u(0.0)), ("dir", u(1.0))])
    cgo = ocg.CgMesh(mo, p)
    Ao, b_o = ocg.cg_stiffness_and_rhs(cgo, mo, func, bo)
    return cgm, A, b, cgo, Ao, b_o


@pytest.mark.parametrize("kind", ["addSchwarz", "hybridSchwarz"])
@pytest.mark.parametrize("n,p", [(16, 4), (8, 2), (5, 1), (12, 8)])
def test_schwarz_operator_is_block_tridiagonal(kind, n, p):
    cgm, A, b, cgo, Ao, b_o = _problem(n, p)
    s = aggmg.cg_smoother(cgm, A, kind)
    so = osm.cg_smoother(cgo, Ao, kind)
    assert np.array_equal(s.mBlockInds, so.mBlockInds)                 # index maps: bit-exact
    slots = blk.level_slots(cgm)
    S = blk.blocks_to_csc(*schwarz_tridiag_blocks(s, slots), slots, A.shape[0]).toarray()
    Y = osm.apply_smoother(so, np.eye(A.shape[0]))
    assert np.abs(S - Y).max() <= 1e-14 * np.abs(Y).max()
    if kind == "hybridSchwarz":
        assert np.array_equal(s.mCountingMatrix, so.mCountingMatrix)


@pytest.mark.gpu
def test_cg_smoother_script():
    """tests/cg_smoother_test.jl:16-48 and :52-104."""
    cgm, A, b, cgo, Ao, b_o = _problem()
    rng = np.random.default_rng(0)
    r = rng.standard_normal(len(b))
    settings = {"jac": 0.5, "addSchwarz": 0.5, "hybridSchwarz": 1.0}      # alpha of the script's solves
    for kind, alpha in settings.items():
        s = aggmg.cg_smoother(cgm, A, kind)
        so = osm.cg_smoother(cgo, Ao, kind)
        y_or = osm.apply_smoother(so, r, alpha=2.0 / 3.0)
        assert np.abs(aggmg.apply_smoother(s, r, alpha=2.0 / 3.0) - y_or).max() <= 1e-12 * np.abs(y_or).max(), kind
        Y_or = osm.apply_smoother(so, Ao, alpha=2.0 / 3.0)               # R = I - apply_smoother(s, A) of the script
        Y = aggmg.apply_smoother(s, A, alpha=2.0 / 3.0)
        assert Y.shape == Y_or.shape and np.abs(Y - Y_or).max() <= 1e-12 * np.abs(Y_or).max(), kind
        x_o, it_o, res_o, err_o = osolv.iterative_smoother_solve(Ao, so, np.zeros(len(b_o)), b_o, maxiter=10 ** 4, alpha=alpha)
        x, it, res, err = aggmg.iterative_smoother_solve(A, s, np.zeros(len(b)), b, maxiter=10 ** 4, alpha=alpha)
        assert it == it_o, (kind, it, it_o)
        assert np.allclose(res, res_o, rtol=1e-8, atol=1e-12 * np.linalg.norm(b_o)), kind
        assert np.abs(x - x_o).max() <= 1e-8 * np.abs(x_o).max(), kind
    # ten damped smoothing steps on sine modes (the script's plots)
    cgm, A, b, cgo, Ao, b_o = _problem(func=lambda x: -np.pi ** 2 * np.sin(np.pi * x), u=lambda x: np.sin(np.pi * x))
    xs = np.zeros(cgm.mNumNodes)
    xs[np.asarray(cgm.mNodesInd).ravel()] = np.asarray(cgm.mNodesX).ravel()
    for kind in settings:
        s = aggmg.cg_smoother(cgm, A, kind)
        so = osm.cg_smoother(cgo, Ao, kind)
        for i in (1, 4, 10):
            u = np.sin(i * np.pi * xs)
            ug, uo = u.copy(), u.copy()
            for _ in range(10):
                ug = ug - aggmg.apply_smoother(s, A @ ug, alpha=2.0 / 3.0)
                uo = uo - osm.apply_smoother(so, Ao @ uo, alpha=2.0 / 3.0)
            assert np.abs(ug - uo).max() <= 1e-11 * max(np.abs(uo).max(), 1.0), (kind, i)
