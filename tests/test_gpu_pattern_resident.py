"""Pattern-resident operators (option pattern_resident = 1 or 2, PatOp / ParamOp in csrc/kernels_fused.cuh).

On a uniform mesh every level is translation invariant: amg1d_set_level_pattern receives n_head + 1 +
n_tail distinct block sets.  By default the device still stores (and the fused legs stream) one block set
per element - the general layout.  With pattern_resident = 1 the fused legs take an element's block set
from the level's small pattern table instead, indexed by the GLOBAL element number.  The numbers and the
order of the arithmetic are the same, so every iterate must be bit-identical to the streamed-operator run
(which in turn is bit-identical to the generic tier and within 1e-10 of the oracle,
tests/test_gpu_parity.py, tests/test_gpu_fullsize.py), the launch count per cycle must not change, and the
residual histories of whole solves must agree exactly (same CTA partition of the norm)."""
import math

import numpy as np
import pytest

from agglomerationmultigrid1d_b200 import uniform

pytestmark = pytest.mark.gpu

SWEEPS = ((3, 3, 2.0 / 3.0), (0, 2, 0.7), (2, 0, 0.5), (1, 1, 1.0), (5, 4, 0.6))


def _check(U, dev, b):
    # the large levels are uploaded as patterns (the few coarsest ones element by element)
    assert all(dev.info(f"pattern:{l}") == (1 if getattr(lv, "is_cg", False) or not lv.explicit else 0)
               for l, lv in enumerate(U.levels))
    assert dev.info("pattern:0") == 1
    assert dev.info("pattern_resident") == 0                       # the general layout is the default
    rng = np.random.default_rng(11)
    x0 = rng.standard_normal(len(b))
    ref = {(a, c): dev.vcycle(x0, b, nPre=a, nPost=c, alpha=al) for a, c, al in SWEEPS}
    x_ref, it_ref, res_ref, _ = dev.solve(np.zeros(len(b)), b, 40, 1e-10)
    y_ref = dev.ldiv(b)
    dev.dev_set_problem(x0, b)
    dev.dev_vcycle(with_residual_norm=True)
    dev.synchronize()
    launches = dev.info("launches_per_cycle")
    # 1: every thread fetches its block set from the table; 2: the CTAs whose window lies in the interior of
    # the level take the interior block set as a by-value kernel parameter (constant-bank operands, ParamOp /
    # f_down_c / f_up_c), the CTAs at the ends of the level keep the table path
    for mode in (1, 2):
        dev.set_option("pattern_resident", mode)
        assert dev.info("pattern_resident") == mode
        for a, c, al in SWEEPS:
            assert np.array_equal(dev.vcycle(x0, b, nPre=a, nPost=c, alpha=al), ref[(a, c)]), (mode, a, c)
        x, it, res, _ = dev.solve(np.zeros(len(b)), b, 40, 1e-10)
        assert it == it_ref and np.array_equal(x, x_ref) and np.array_equal(res, res_ref), mode
        assert np.array_equal(dev.ldiv(b), y_ref), mode
        dev.dev_set_problem(x0, b)
        dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        assert dev.info("launches_per_cycle") == launches
    with pytest.raises(Exception):
        dev.set_option("pattern_resident", 3)
    # the streamed-operator byte model drops the operator of every pattern level
    assert U.bytes_per_leg_fused(0, down=False) == 8 * (U.levels[0].n * 3 * U.levels[0].m
                                                         + U.levels[1].n * U.levels[1].m)
    dev.set_option("pattern_resident", 0)
    assert U.bytes_per_leg_fused(0, down=False) > 8 * U.levels[0].n * (3 * U.levels[0].m + U.levels[0].m ** 2)
    assert np.array_equal(dev.vcycle(x0, b), ref[(3, 3)])


# element counts that are not a multiple of a CTA's output window (120 / 56 elements); 3 * 2^k: the
# agglomeration stops at 3 elements
@pytest.mark.parametrize("orders,n", [((3, 1), 1536), ((3, 1), 196608), ((4, 2, 1), 98304), ((2, 1), 3072),
                                      ((1,), 49152), ((8, 4, 2, 1), 3072)])
def test_dg_first_hierarchy(orders, n):
    k = (n & -n).bit_length() - 1
    U = uniform.UniformDgHierarchy(n, list(orders), [2] * k, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    try:
        w = 2.0 * math.pi / 64.0
        b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
        _check(U, dev, b)
    finally:
        dev.close()


@pytest.mark.parametrize("cg,dg,n", [((3, 1), (1,), 196608), ((3, 1), (1,), 1536), ((4, 2, 1), (), 3072),
                                     ((1,), (), 49152)])
def test_cg_first_hierarchy(cg, dg, n):
    """CG levels in group form: n + 1 groups, point Jacobi, two-parent transfers (BASELINE C4 shape)."""
    k = (n & -n).bit_length() - 1
    agg = [2] * k if dg else [4] + [2] * (k - 2)
    U = uniform.UniformCgHierarchy(n, list(cg), list(dg), agg, xin=0.0, xout=1.0, CDir=1000.0 * n)
    dev = U.upload()
    try:
        b = U.rhs(np.cos, [-math.sin(0.0), math.cos(1.0)])
        _check(U, dev, b)
    finally:
        dev.close()
