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
