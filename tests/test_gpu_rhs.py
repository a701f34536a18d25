"""Device-side right-hand side assembly against the host assemblies:
the product's vectorised one (uniform.py / dg_mesh.py) and the oracle's literal element loops
(oracle/dg.py: dg_flux_rhs, oracle/cg.py: cg_stiffness_and_rhs), to 1e-13 of max|b|; and multigrid on the
device-resident problem (amg1d_dev_solve) against multigrid with the host-assembled vector."""
import math

import numpy as np
import pytest

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import uniform
from agglomerationmultigrid1d_b200.reference_element import ReferenceElement

pytestmark = pytest.mark.gpu
W = 2.0 * math.pi / 64.0


def test_dg_rhs_on_device_matches_host_and_oracle():
    from oracle import drivers
    n = 1024
    U = uniform.UniformDgHierarchy(n, [3, 1], [2] * 10, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    bc = [0.0, math.cos(W * n)]
    U.device_rhs([("cos", W * W, 0, W, 0.0)], bc)
    b_dev = dev.dev_get_rhs()
    b_host = U.rhs(lambda x: W * W * np.cos(W * x), bc)
    _, _, b_or, _ = drivers.dg_agg_problem(n, p=3, unit_h=True)
    scale = np.abs(b_or).max()
    assert np.abs(b_dev - b_host).max() <= 1e-13 * scale
    assert np.abs(b_dev - b_or).max() <= 1e-13 * scale
    # multigrid on the device-resident problem: same count, same history as with the host vector
    it_d, res_d = dev.dev_solve(100, 1e-10)
    x_d = dev.dev_get_solution()
    x_h, it_h, res_h, _ = dev.solve(np.zeros(len(b_host)), b_host, 100, 1e-10)
    assert it_d == it_h == 11
    assert np.all(np.abs(res_d - res_h) <= np.maximum(1e-10 * res_h, 1e-13 * np.linalg.norm(b_host)))
    assert np.abs(x_d - x_h).max() <= 1e-9 * np.abs(x_h).max()
    dev.close()


def test_function_families():
    """sum of terms coef * x^pow * g(w x + phi): constants, powers, cos / sin / exp, on [0, 1] (the reference's
    scripts use cos(x), exp(-x) and 1)."""
    n = 256
    U = uniform.UniformDgHierarchy(n, [3, 1], [2] * 8, xin=0.0, xout=1.0, CDir=1000.0 * n, bc_kinds=("dir", "dir"))
    dev = U.upload()
    terms = [("1", 2.0, 0, 0.0, 0.0), ("1", 3.0, 2, 0.0, 0.0), ("exp", 1.0, 0, -1.0, 0.0), ("sin", -0.5, 1, 2.0, 0.3)]
    f = lambda x: 2.0 + 3.0 * x ** 2 + np.exp(-x) - 0.5 * x * np.sin(2.0 * x + 0.3)      # noqa: E731
    bc = [0.7, -1.3]
    U.device_rhs(terms, bc)
    b_dev, b_host = dev.dev_get_rhs(), U.rhs(f, bc)
    assert np.abs(b_dev - b_host).max() <= 1e-13 * np.abs(b_host).max()
    dev.close()


@pytest.mark.parametrize("cg", [[3, 1], [1], [8, 4, 2, 1]])
def test_cg_rhs_on_device_matches_host_and_oracle(cg):
    from oracle import drivers
    n = 256
    dg, agg = ([1], [2] * 8) if cg != [1] else ([], [2])
    if cg[0] == 8:
        dg, agg = [], [4] + [2] * 6
    U = uniform.UniformCgHierarchy(n, cg, dg, agg, xin=0.0, xout=1.0, CDir=1000.0 * n)
    dev = U.upload()
    bc = [-math.sin(0.0), math.cos(1.0)]
    U.device_rhs([("cos", 1.0, 0, 1.0, 0.0)], bc)
    b_dev = dev.dev_get_rhs()
    b_host = U.rhs(np.cos, bc)
    _, _, b_or, _ = drivers.build_problem(n, cg_orders=cg, dg_orders=dg, agg_factors=agg)
    s0 = U.group_slots(0).ravel()
    ok = s0 >= 0
    scale = np.abs(b_or).max()
    assert np.abs(b_dev - b_host).max() <= 1e-13 * scale
    assert np.abs(b_dev[ok] - b_or[s0[ok]]).max() <= 1e-13 * scale
    assert np.all(b_dev[~ok] == 0.0)                    # padding slots of the closing group
    it_d, res_d = dev.dev_solve(100, 1e-10)
    x_h, it_h, res_h, _ = dev.solve(np.zeros(len(b_host)), b_host, 100, 1e-10)
    assert it_d == it_h
    dev.close()


def test_volume_integrals_on_a_nonuniform_mesh():
    """The general (vertex array) form against the product's element-by-element host assembly on a graded mesh:
    with homogeneous boundary data dg_flux_rhs returns the pure volume integrals (src/dg_mesh.jl:342-365)."""
    n = 2 ** 12
    rng = np.random.default_rng(0)
    x = np.concatenate([[0.0], np.cumsum(0.5 + rng.random(n))])
    mesh = aggmg.Mesh(x)
    bd = aggmg.set_boundary(mesh, 0.0, float(x[-1]), [("neu", 0.0), ("dir", 0.0)])
    dgm = aggmg.DgMesh(mesh, 3)
    f, r = aggmg.dg_flux_rhs(dgm, mesh, lambda t: W * W * np.cos(W * t), bd, 1000.0)
    assert np.abs(r).max() == 0.0
    G, D, C = aggmg.dg_flux_operators(dgm, mesh, bd, 1000.0)
    A = (C - D @ dgm.mMassMatrixLU.solve(G)).tocsc()
    H = aggmg.MeshHierarchy([dgm, aggmg.DgMesh(mesh, 1)], [bd, bd], A, G, D, C, nDG=2)
    ref = ReferenceElement(3)
    Wq = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal
    H.device.dev_assemble_rhs(0, ref.mGaussQuadNodes, Wq, [("cos", W * W, 0, W, 0.0)], 0.0, float(x[-1]), vertices=x)
    b_dev = H.device.dev_get_rhs()
    assert np.abs(b_dev - f).max() <= 1e-13 * np.abs(f).max()
    H.device.close()


def test_argument_checks():
    from agglomerationmultigrid1d_b200 import _capi as capi
    n = 64
    U = uniform.UniformDgHierarchy(n, [3, 1], [2] * 6, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    ref = ReferenceElement(3)
    Wq = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal
    with pytest.raises(capi.Amg1dError):                 # wrong basis size for level 0
        dev.dev_assemble_rhs(0, ref.mGaussQuadNodes, Wq[:, :3], [("cos", 1.0, 0, 1.0, 0.0)], 0.0, 1.0)
    with pytest.raises(capi.Amg1dError):                 # unknown function family
        dev.dev_assemble_rhs(0, ref.mGaussQuadNodes, Wq, [(7, 1.0, 0, 1.0, 0.0)], 0.0, 1.0)
    with pytest.raises(capi.Amg1dError):                 # CG kind on a DG level
        dev.dev_assemble_rhs(1, ref.mGaussQuadNodes, Wq, [("cos", 1.0, 0, 1.0, 0.0)], 0.0, 1.0)
    dev.close()
