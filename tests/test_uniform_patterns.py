"""The scalable uniform-mesh set-up (head / interior / tail patterns, closed-form transfer blocks)
must reproduce the general path (the reference's sparse algebra) block by block."""
import math

import numpy as np
import pytest

from agglomerationmultigrid1d_b200 import blocks as blk
from agglomerationmultigrid1d_b200 import uniform
from shapes import build_package

CASES = [
    (128, [3, 1], [2] * 7, 1, False),
    (256, [4, 2, 1], [2] * 8, 1, True),
    (192, [2, 1], [3, 2, 2, 2, 2, 2, 2], 1, False),
    (128, [1], [4, 2, 2, 2, 2, 2], 0, False),
    (128, [3, 1, 0], [], 1, False),
    (16, [2, 1], [2, 2], 1, False),          # everything explicit
]


@pytest.mark.parametrize("n,orders,agg,pAgg,unit_h", CASES)
def test_patterns_match_general_path(n, orders, agg, pAgg, unit_h):
    Hp, _, bp = build_package(n, dg_orders=orders, agg_factors=agg, pAgg=pAgg, unit_h=unit_h, upload=False)
    if unit_h:
        w = 2 * math.pi / 64
        kw = dict(xin=0.0, xout=float(n), CDir=1000.0)
        func, vals = (lambda x: w * w * np.cos(w * x)), [-w * math.sin(0.0), math.cos(w * n)]
    else:
        kw = dict(xin=0.0, xout=1.0, CDir=1000.0 * n)
        func, vals = np.cos, [-math.sin(0.0), math.cos(1.0)]
    U = uniform.UniformDgHierarchy(n, orders, agg, pAgg=pAgg, **kw)
    assert len(U.levels) == len(Hp.mMeshes)
    for l, lv in enumerate(U.levels):
        lo, di, up = U.level_blocks(l)
        glo, gdi, gup = blk.csc_to_blocks(Hp.mStiffness[l], blk.level_slots(Hp.mMeshes[l]))
        scale = np.abs(gdi).max()
        for a, g in ((lo, glo), (di, gdi), (up, gup)):
            assert np.abs(a - g).max() <= 1e-13 * scale, (l, lv.n)
    for l, (P, ratio) in enumerate(U.transfers):
        L = uniform._transfer_csc(P, U.levels[l].n, ratio)
        assert abs(L - Hp.mInterpolation[l]).max() <= 1e-13
    b = U.rhs(func, vals)
    assert np.abs(b - bp).max() <= 1e-13 * np.abs(bp).max()
    # slabs of the right-hand side concatenate to the whole vector (multi-GPU assembly)
    q = n // 4
    parts = [U.rhs(func, vals, elem_range=(i * q, (i + 1) * q)) for i in range(4)]
    assert np.array_equal(np.concatenate(parts), b)


def test_byte_models():
    """B_ref reproduces SURVEY 8d (T: 493.9 GB, C2: 7.72 GB per V-cycle with check)."""
    U = uniform.UniformDgHierarchy(2 ** 26, [3, 1], [2] * 26, xin=0.0, xout=float(2 ** 26), CDir=1000.0)
    assert abs(U.bytes_per_cycle_reference_model() / 1e9 - 493.9) < 0.1
    assert U.dof_updates_per_cycle() == 6 * (2 ** 28 + 2 ** 27 + sum(2 ** k * 2 for k in range(1, 26)))
    U2 = uniform.UniformDgHierarchy(2 ** 20, [3, 1], [2] * 20, xin=0.0, xout=float(2 ** 20), CDir=1000.0)
    assert abs(U2.bytes_per_cycle_reference_model() / 1e9 - 7.72) < 0.01
    assert U.bytes_per_cycle_fused() < U.bytes_per_cycle_reference_model() / 3.5


CG_CASES = [
    (256, [3, 1], [1], [2] * 8, 1),                       # BASELINE C4 shape
    (256, [8, 4, 2, 1], [], [4, 2, 2, 2, 2, 2, 2], 1),    # tests/full_heirarchy_test.jl
    (128, [1], [], [2], 1),                               # BASELINE C1
    (128, [8, 4, 2, 1], [0], [], 1),                      # tests/dg_cg_heirarchy_test.jl
    (128, [2, 1], [], [], 1),                             # tests/cg_heirarchy_test.jl (two levels)
]


@pytest.mark.parametrize("n,cg,dg,agg,pAgg", CG_CASES)
def test_cg_patterns_match_general_path(n, cg, dg, agg, pAgg):
    import scipy.sparse as sp
    Hp, _, bp = build_package(n, cg_orders=cg, dg_orders=dg, agg_factors=agg, pAgg=pAgg, upload=False)
    U = uniform.UniformCgHierarchy(n, cg, dg, agg, pAgg=pAgg, xin=0.0, xout=1.0, CDir=1000.0 * n)
    assert len(U.levels) == len(Hp.mMeshes)
    for l in range(len(U.levels)):
        lo, di, up = U.level_blocks(l)
        glo, gdi, gup = blk.csc_to_blocks(Hp.mStiffness[l], blk.level_slots(Hp.mMeshes[l]))
        for a, g in ((lo, glo), (di, gdi), (up, gup)):
            assert np.abs(a - g).max() <= 1e-13 * np.abs(gdi).max(), l
    for l in range(len(U.levels) - 1):
        parent, P0, P1 = U.transfer_blocks(l)
        assert np.all(np.diff(parent) >= 0) and parent[0] >= -1
        sf, sc = blk.level_slots(Hp.mMeshes[l]), blk.level_slots(Hp.mMeshes[l + 1])
        L = sp.csc_matrix(Hp.mInterpolation[l])
        xh = np.random.default_rng(l).standard_normal(L.shape[1])
        nc, mc = sc.shape
        xc = np.zeros((nc + 2, mc))
        xc[1:-1][sc >= 0] = xh[sc[sc >= 0]]
        xf = np.einsum("eij,ej->ei", P0, xc[parent + 1])
        if P1 is not None:
            xf += np.einsum("eij,ej->ei", P1, xc[parent + 2])
        ref = L @ xh
        assert np.abs(xf[sf >= 0] - ref[sf[sf >= 0]]).max() <= 1e-13 * max(1.0, np.abs(ref).max()), l
    b = U.rhs(np.cos, [-math.sin(0.0), math.cos(1.0)])
    s0 = U.group_slots(0)
    assert np.array_equal(s0, blk.level_slots(Hp.mMeshes[0]))
    bg = np.zeros(s0.shape)
    bg[s0 >= 0] = bp[s0[s0 >= 0]]
    assert np.abs(b - bg.ravel()).max() <= 1e-13 * np.abs(bp).max()
