"""GPU tests at BASELINE sizes.  The oracle cannot run there, so parity is carried by
(a) the pattern upload path checked against the oracle at n = 1024, and
(b) size-independent properties at 2^20 and 2^26 elements: exact linearity in powers of two, agreement of
the fused residual norm with an independently computed one, symmetry of the operator, bit-identical
results of the two kernel tiers, monotone convergence with the mesh-independent cycle count."""
import math

import numpy as np
import pytest

from agglomerationmultigrid1d_b200 import uniform

pytestmark = pytest.mark.gpu


def _problem(n):
    w = 2.0 * math.pi / 64.0
    return (lambda x: w * w * np.cos(w * x)), [0.0, math.cos(w * n)]


def _build(log2n, orders=(3, 1)):
    n = 2 ** log2n
    U = uniform.UniformDgHierarchy(n, list(orders), [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
    return U, U.upload()


def test_pattern_path_matches_oracle_n1024():
    from oracle import drivers, solvers
    n = 1024
    U, dev = _build(10)
    func, vals = _problem(n)
    b = U.rhs(func, vals)
    H, x0, bo, _ = drivers.dg_agg_problem(n, p=3, unit_h=True)
    assert np.abs(b - bo).max() <= 1e-13 * np.abs(bo).max()
    x_or, it_or, res_or, _ = solvers.multigrid(H, x0, bo, 100, 1e-10, u_exact=np.zeros(len(bo)))
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it == it_or == 11
    assert np.all(np.abs(res - res_or) <= np.maximum(1e-10 * res_or, 1e-13 * np.linalg.norm(bo)))
    assert np.abs(x - x_or).max() <= 1e-9 * np.abs(x_or).max()
    dev.close()


@pytest.mark.parametrize("log2n", [20, 26])
def test_properties_at_baseline_sizes(log2n):
    n = 2 ** log2n
    U, dev = _build(log2n)
    func, vals = _problem(n)
    b = U.rhs(func, vals)
    x0 = np.zeros(len(b))
    # convergence: monotone, mesh-independent count (11 at n <= 2^13, slowly growing: 13 at 2^20, 15 at 2^26)
    x, it, res, _ = dev.solve(x0, b, 100, 1e-10)
    assert it <= 16 and np.all(np.diff(res) < 0)
    assert res[-1] < 1e-10 * np.linalg.norm(b)
    assert np.all(res[1:6] / res[:5] < 0.35)            # ~0.2 reduction per cycle
    # the fused residual norm (inside f_up) equals the one from the stand-alone residual kernel
    r = dev.residual(0, x, b)
    assert abs(np.linalg.norm(r) - res[-1]) <= 1e-12 * np.linalg.norm(b) + 1e-9 * res[-1]
    # exact linearity under scaling by a power of two (no rounding involved)
    x1 = dev.vcycle(x0, b)
    x2 = dev.vcycle(x0, 4.0 * b)
    assert np.array_equal(x2, 4.0 * x1)
    # the single-CTA coarse tail replaces ~20 small launches without changing a bit
    assert dev.info("tail_start") > 0
    dev.set_option("coarse_cta_elems", 0)
    assert dev.info("tail_start") < 0
    assert np.array_equal(dev.vcycle(x0, b), x1)
    dev.set_option("coarse_cta_elems", 1024)
    assert dev.info("structure:0") == 1 and dev.info("tile_rows:0") == 40      # 4 + 16 + 4 + 16 doubles
    # every level above the tail runs as exactly one fused kernel per leg (a silent fall-back to the
    # streaming kernels would show up here as 4-5 launches per level)
    dev.dev_vcycle(with_residual_norm=True)
    dev.synchronize()
    assert dev.info("launches_per_cycle") <= 2 * dev.info("tail_start") + 1 + 3
    # both kernel tiers give bit-identical iterates (generic tier only at 2^20: it is ~4x slower)
    if log2n <= 20:
        dev.set_option("fused", 0)
        xg = dev.vcycle(x0, b)
        dev.set_option("fused", 1)
        assert np.array_equal(xg, x1)
        # symmetry of the fine operator: u'(A v) = v'(A u)
        rng = np.random.default_rng(0)
        u, v = rng.standard_normal(len(b)), rng.standard_normal(len(b))
        s1, s2 = float(u @ dev.matvec(0, v)), float(v @ dev.matvec(0, u))
        assert abs(s1 - s2) <= 1e-9 * (np.linalg.norm(u) * np.linalg.norm(v) * 1000.0)
    dev.close()


@pytest.mark.parametrize("log2n", [18, 24])
def test_p4_hierarchy_properties(log2n):
    """BASELINE C3 shape (DG p=4 -> 2 -> 1 -> agglomerated) at 2^18 and at its full 2^24 elements:
    fused kernels with 5x5 blocks against the generic tier."""
    U, dev = _build(log2n, orders=(4, 2, 1))
    func, vals = _problem(2 ** log2n)
    b = U.rhs(func, vals)
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it <= 16 and np.all(np.diff(res) < 0) and res[-1] < 1e-10 * np.linalg.norm(b)
    dev.set_option("fused", 0)
    xg, itg, resg, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert itg == it and np.array_equal(xg, x)
    dev.close()


@pytest.mark.parametrize("cg,dg,agg", [([3, 1], [1], [2] * 8), ([8, 4, 2, 1], [], [4] + [2] * 6), ([1], [], [2])])
def test_cg_pattern_path_matches_oracle(cg, dg, agg):
    """CG-first hierarchies through the pattern upload path (two-parent pattern transfers, point
    Jacobi, group-ordered vectors) against the oracle at n = 256 (C4, full_heirarchy, C1 shapes)."""
    import scipy.sparse as sp
    from oracle import drivers, solvers
    from test_gpu_parity import dense_coarse_solver
    n = 256
    H, x0, bo, _ = drivers.build_problem(n, cg_orders=cg, dg_orders=dg, agg_factors=agg)
    U = uniform.UniformCgHierarchy(n, cg, dg, agg, xin=0.0, xout=1.0, CDir=1000.0 * n)
    dev = U.upload()
    b = U.rhs(np.cos, [-math.sin(0.0), math.cos(1.0)])
    s0 = U.group_slots(0)
    valid = s0.ravel() >= 0
    x_or, it_or, res_or, _ = solvers.multigrid(H, x0, bo, 100, 1e-10)
    with dense_coarse_solver():
        _, it_d, res_d, _ = solvers.multigrid(H, x0, bo, 100, 1e-10)
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it == it_or
    k = min(it_d, it_or)
    noise = np.zeros(it_or)
    noise[:k] = 20 * np.abs(res_d[:k] - res_or[:k])
    A = sp.csr_matrix(H.mStiffness[0])
    floor = 64 * np.finfo(float).eps * abs(A).sum(axis=1).max() * np.abs(x_or).max() * np.sqrt(A.shape[0])
    assert np.all(np.abs(res - res_or) <= np.maximum(np.maximum(1e-10 * res_or, floor), noise))
    xh = np.zeros(len(bo))
    xh[s0.ravel()[valid]] = x[valid]
    assert np.abs(xh - x_or).max() <= 1e-8 * np.abs(x_or).max()
    assert np.all(x[~valid] == 0.0)
    # CG levels run in the fused point-Jacobi kernels with two-parent transfers: same bits as the
    # generic tier, with and without the compressed row / column structure
    if cg[0] >= 2:
        assert dev.info("structure:0") == 2
    dev.set_option("fused", 0)
    xg, itg, _, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    dev.set_option("fused", 1)
    assert itg == it and np.array_equal(xg, x)
    dense = U.upload(options={"compress": 0})
    xd, itd, _, _ = dense.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert itd == it and np.array_equal(xd, x)
    rng = np.random.default_rng(2)
    x0 = rng.standard_normal(len(b))
    x0[~valid] = 0.0
    for nPre, nPost in ((3, 3), (0, 2), (2, 0), (1, 4)):
        ref = None
        for d_, fused in ((dev, 1), (dev, 0), (dense, 1)):
            d_.set_option("fused", fused)
            got = d_.vcycle(x0, b, nPre=nPre, nPost=nPost)
            ref = got if ref is None else ref
            assert np.array_equal(got, ref), (nPre, nPost, fused)
    dev.set_option("fused", 1)
    dense.close()
    dev.close()


def test_c4_shape_at_scale():
    """BASELINE C4 (CG 3 -> 1 -> DG 1 -> agglomerated) at 2^22 elements: convergence properties."""
    n = 2 ** 22
    w = 2.0 * math.pi / 64.0
    U = uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * 22, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it <= 20 and np.all(np.diff(res) < 0) and res[-1] < 1e-10 * np.linalg.norm(b)
    x1 = dev.vcycle(np.zeros(len(b)), b)
    assert np.array_equal(dev.vcycle(np.zeros(len(b)), 2.0 * b), 2.0 * x1)
    assert dev.info("launches_per_cycle") <= 2 * dev.info("tail_start") + 1      # CG levels fused too
    dev.set_option("fused", 0)
    assert np.array_equal(dev.vcycle(np.zeros(len(b)), b), x1)       # generic tier, same bits
    dev.close()


def test_direct_solver_at_scale():
    """Block cyclic reduction at sizes no host direct solver is asked to do here:
    A \\ b on the finest DG p=3 level of 2^20 elements (4 M unknowns), and the literal dg_cg_heirarchy
    shape (CG 3 -> 1 -> DG 0, src/mesh_heirarchy.jl:57-74) at 2^22 elements, whose coarsest level is a
    2^22-unknown tridiagonal system solved exactly inside every V-cycle."""
    n = 2 ** 20
    U, dev = _build(20)
    func, vals = _problem(n)
    b = U.rhs(func, vals)
    x = dev.direct_solve(0, b)
    r = dev.residual(0, x, b)
    assert np.linalg.norm(r) <= 1e-6 * np.linalg.norm(b)        # cond ~ 1e12 at this size: backward stable
    xm, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert np.abs(x - xm).max() <= 1e-5 * np.abs(xm).max()
    dev.close()
    n = 2 ** 22
    w = 2.0 * math.pi / 64.0
    U = uniform.UniformCgHierarchy(n, [3, 1], [0], [], xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    assert dev.info("n_levels") == 3 and dev.info("tail_start") < 0
    b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it <= 25 and res[-1] < 1e-10 * np.linalg.norm(b) and np.all(np.diff(res) < 0)
    dev.close()


def test_device_side_setup_nonuniform_mesh():
    """Device-side set-up on a graded, non-uniform mesh of 2^14 elements - the case the
    uniform-mesh pattern path cannot take: every level's blocks against the host's sparse algebra, and the
    solve converges at the usual rate."""
    import time
    import agglomerationmultigrid1d_b200 as aggmg
    n = 2 ** 14
    rng = np.random.default_rng(0)
    hs = 0.5 + rng.random(n)                                   # element sizes within a factor of 3
    x = np.concatenate([[0.0], np.cumsum(hs)])
    xout = float(x[-1])
    mesh = aggmg.Mesh(x)
    w = 2.0 * math.pi / 64.0
    bd = aggmg.set_boundary(mesh, 0.0, xout, [("neu", 0.0), ("dir", math.cos(w * xout))])
    meshes = [aggmg.DgMesh(mesh, 3), aggmg.DgMesh(mesh, 1)]
    cur = n
    for i in range(10):
        agg = [[2 * j, 2 * j + 1] for j in range(cur // 2)]
        cur //= 2
        meshes.append(aggmg.AgglomeratedDgMesh1(1, agg, mesh, meshes[1]) if i == 0
                      else aggmg.AgglomeratedDgMeshN(1, agg, meshes[-1], meshes[1]))
    G, D, C = aggmg.dg_flux_operators(meshes[0], mesh, bd, 1000.0)
    A = (C - D @ meshes[0].mMassMatrixLU.solve(G)).tocsc()
    f, r = aggmg.dg_flux_rhs(meshes[0], mesh, lambda t: w * w * np.cos(w * t), bd, 1000.0)
    b = f - D @ meshes[0].mMassMatrixLU.solve(r)
    t0 = time.perf_counter()
    Hh = aggmg.MeshHierarchy(meshes, [bd] * len(meshes), A, G, D, C, nDG=2, nAgg=10)
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    Hd = aggmg.MeshHierarchy(meshes, [bd] * len(meshes), A, G, D, C, nDG=2, nAgg=10, device_setup=True)
    t_dev = time.perf_counter() - t0
    print(f"set-up of 12 levels, 2^14 non-uniform elements: host {t_host:.2f} s, device-side {t_dev:.2f} s")
    for l in range(len(meshes)):
        got, ref = Hd.level_blocks(l), Hh.level_blocks(l)
        scale = np.abs(ref[1]).max()
        assert all(np.abs(g - r_).max() <= 1e-12 * scale for g, r_ in zip(got[:3], ref[:3])), l
    xh, ith, resh, _ = aggmg.multigrid(Hh, np.zeros(len(b)), b, 100, 1e-10, with_error=False)
    xd, itd, resd, _ = aggmg.multigrid(Hd, np.zeros(len(b)), b, 100, 1e-10, with_error=False)
    assert itd == ith and ith <= 20 and resd[-1] < 1e-10 * np.linalg.norm(b)
    assert np.allclose(resd, resh, rtol=1e-8, atol=1e-13 * np.linalg.norm(b))
    # both iterates have residuals of 1e-10 ||b||; the operators of the two set-ups differ by 1e-12 relative and
    # cond(A) ~ 1e8 on this mesh, so the iterates themselves agree to ~1e-8
    assert np.abs(xd - xh).max() <= 1e-7 * np.abs(xh).max()
    Hh.device.close()
    Hd.device.close()
