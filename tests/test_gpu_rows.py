"""Row-per-thread fused legs (csrc/kernels_rows.cuh) for the large element blocks of the reference's own
hierarchy scripts (DG / CG p = 8 -> m = 9 / 8, tests/dg_heirarchy_test.jl, tests/dg_cg_heirarchy_test.jl)
and the orders in between.

They must reproduce the generic tier bit for bit (same accumulation order per block row), for both window
sizes and 1, 2 or 3 block rows per thread, at element counts that are not a multiple of a CTA's output window, and for sweep counts that
change the halo.  The oracle comparison of the same shapes is in tests/test_gpu_parity.py (n = 32) and
tests/test_gpu_fullsize.py::test_cg_pattern_path_matches_oracle (n = 256); here the sizes are large
enough for several thousand CTAs."""
import math

import numpy as np
import pytest

from agglomerationmultigrid1d_b200 import uniform

pytestmark = pytest.mark.gpu

ROWS_PER_THREAD_DEFAULT = 0    # amg1d.cu: opt_rows_rpt, 0 = auto (1 streamed, 3 pattern-resident)


def _rhs_dg(U, n):
    w = 2.0 * math.pi / 64.0
    return U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])


def _levels_with_big_blocks(U):
    return [l for l, lv in enumerate(U.levels[:-1]) if lv.m > 5 or (lv.m == 5 and U.levels[l + 1].m == 2)]


def _compare_tiers(dev, b, n_big_levels):
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(len(b))
    # reference: the generic tier (one kernel per operation, any block size)
    dev.set_option("fused", 0)
    ref = {}
    sweeps = ((3, 3, 2.0 / 3.0), (0, 2, 0.7), (2, 0, 0.5), (1, 1, 1.0), (5, 4, 0.6))
    for nPre, nPost, alpha in sweeps:
        ref[(nPre, nPost)] = dev.vcycle(x0, b, nPre=nPre, nPost=nPost, alpha=alpha)
    x_ref, it_ref, res_ref, _ = dev.solve(np.zeros(len(b)), b, 40, 1e-10)
    dev.set_option("fused", 1)
    launches = {}
    for window, rpt in ((0, 1), (32, 1), (64, 1), (32, 2), (64, 2), (32, 3), (64, 3)):
        dev.set_option("rows_window", window)
        dev.set_option("rows_per_thread", rpt)
        for nPre, nPost, alpha in sweeps:
            got = dev.vcycle(x0, b, nPre=nPre, nPost=nPost, alpha=alpha)
            assert np.array_equal(got, ref[(nPre, nPost)]), (window, rpt, nPre, nPost)
        x, it, res, _ = dev.solve(np.zeros(len(b)), b, 40, 1e-10)
        assert it == it_ref and np.array_equal(x, x_ref), (window, rpt)
        # the residual norm is summed in a different order by each tier
        assert np.allclose(res, res_ref, rtol=1e-9, atol=1e-13 * np.linalg.norm(b)), (window, rpt)
        dev.dev_set_problem(x0, b)
        dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        launches[(window, rpt)] = dev.info("launches_per_cycle")
    dev.set_option("rows_window", 64)
    dev.set_option("rows_per_thread", ROWS_PER_THREAD_DEFAULT)
    # one kernel per leg instead of nPre + 1 (down) and nPost + 1 (+ 1 for the norm) streaming passes
    assert len({v for k, v in launches.items() if k[0]}) == 1, launches
    assert launches[(0, 1)] - launches[(64, 1)] >= 5 * n_big_levels, launches
    return launches


# 98304 elements: 1756 (W = 64) / 4096 (W = 32) CTAs per leg
@pytest.mark.parametrize("orders,n", [((8, 4, 2, 1), 3072), ((8, 4, 2, 1), 98304), ((8,), 3072), ((7, 3, 1), 1536),
                                      ((7, 3, 1), 98304), ((6, 3, 1), 1536), ((5, 2, 1), 98304), ((4,), 3072)])
def test_dg_large_blocks_bit_identical(orders, n):
    k = (n & -n).bit_length() - 1                      # n = odd * 2^k: agglomerate down to `odd` elements
    U = uniform.UniformDgHierarchy(n, list(orders), [2] * k, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    try:
        big = _levels_with_big_blocks(U)
        assert big and big[0] == 0
        assert dev.info("structure:0") == 1            # assembled DG level: column / row structure
        _compare_tiers(dev, _rhs_dg(U, n), len(big))
    finally:
        dev.close()


def test_dg_large_blocks_dense_storage():
    """compress = 0 keeps dense off-diagonal blocks: the rows kernels then hold three block rows in registers."""
    n = 1536
    U = uniform.UniformDgHierarchy(n, [8, 4, 2, 1], [2] * 9, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload(options={"compress": 0})
    try:
        assert dev.info("structure:0") == 0
        _compare_tiers(dev, _rhs_dg(U, n), 1)
    finally:
        dev.close()


@pytest.mark.parametrize("cg,n", [((8, 4, 2, 1), 3072), ((8, 4, 2, 1), 98304), ((7, 3, 1), 1536), ((6, 3, 1), 98304),
                                  ((5, 2, 1), 1536)])
def test_cg_large_groups_bit_identical(cg, n):
    """CG levels in group form [vertex_k, interior nodes of element k]: m = p, point Jacobi, two-parent
    transfers (cg_cg_interpolation)."""
    k = (n & -n).bit_length() - 1
    U = uniform.UniformCgHierarchy(n, list(cg), [], [4] + [2] * (k - 2), xin=0.0, xout=1.0, CDir=1000.0 * n)
    dev = U.upload()
    try:
        assert dev.info("structure:0") == 2            # CG groups: row / column structure
        b = U.rhs(np.cos, [-math.sin(0.0), math.cos(1.0)])
        _compare_tiers(dev, b, 1)
    finally:
        dev.close()
