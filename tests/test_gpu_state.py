"""State handling of the device-resident path and the Krylov wrapper (round-1 advisor findings):
  * amg1d_dev_rhs_norm must not disturb the residual norm cached by a norm-fused V-cycle;
  * the per-level host operations that stage through level 0's vectors mark the device-resident problem
    stale instead of letting amg1d_dev_vcycle solve for a wrong right-hand side;
  * amg1d_pcg with b = 0 or an exact initial guess returns x0 with iters = 0, never NaNs with AMG1D_OK."""
import numpy as np
import pytest

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import _capi as capi
from shapes import SHAPES, build_package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prob():
    Hp, _, bp = build_package(**SHAPES["C2_dg3_agg"])
    yield Hp, bp
    Hp.device.close()


def test_rhs_norm_keeps_the_cached_residual_norm(prob):
    Hp, b = prob
    dev = Hp.device
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal(len(b))
    dev.dev_set_problem(x0, b)
    dev.dev_vcycle(with_residual_norm=True)            # ||b - A x|| cached on the device
    nb = dev.dev_rhs_norm()                            # must not overwrite that cache
    res = dev.dev_residual_norm()
    x = dev.dev_get_solution()
    r = dev.residual(0, x, b)
    assert abs(nb - np.linalg.norm(b)) <= 1e-13 * np.linalg.norm(b)
    assert abs(res - np.linalg.norm(r)) <= 1e-12 * np.linalg.norm(r)
    assert res < 0.5 * nb                               # i.e. not ||b|| read back as the residual


def test_stale_device_problem_is_refused(prob):
    Hp, b = prob
    dev = Hp.device
    dev.dev_set_problem(None, b)
    dev.dev_vcycle()
    x1 = dev.dev_get_solution()
    dev.residual(0, np.ones(len(b)), 2.0 * b)           # stages 2 b in level 0's rhs buffer
    with pytest.raises(capi.Amg1dError) as ei:
        dev.dev_vcycle()
    assert ei.value.code == capi.ERR_STATE
    with pytest.raises(capi.Amg1dError):
        dev.dev_rhs_norm()
    with pytest.raises(capi.Amg1dError):
        dev.dev_set_problem(np.zeros(len(b)), None)     # "keep b" cannot be honoured either
    dev.dev_set_problem(None, b)                        # a fresh problem makes the path usable again
    dev.dev_vcycle()
    assert np.array_equal(dev.dev_get_solution(), x1)
    dev.matvec(0, np.ones(len(b)))                      # does not touch b: the problem stays valid
    dev.restrict(0, np.ones(len(b)))
    dev.dev_vcycle()
    dev.pcg(np.zeros(len(b)), b, 5, 1e-10)              # level 0's rhs buffer becomes the CG residual
    with pytest.raises(capi.Amg1dError):
        dev.dev_vcycle()
    # the host-vector entry points upload their own problem and are never affected
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 100, 1e-10)
    assert it > 0 and res[-1] < 1e-10 * np.linalg.norm(b)
    dev.dev_vcycle()


def test_pcg_breakdown_guards(prob):
    Hp, b = prob
    dev = Hp.device
    x0 = np.arange(len(b), dtype=float)
    x, it, res = dev.pcg(x0, np.zeros(len(b)), 50, 1e-10)          # b = 0: nothing to do
    assert it == 0 and len(res) == 0 and np.array_equal(x, x0)
    xs, its, ress = dev.pcg(np.zeros(len(b)), b, 100, 1e-13)
    assert its > 0 and np.all(np.isfinite(xs))
    x, it, res = dev.pcg(xs, b, 50, 1e-9)                          # x0 already solves to the tolerance
    assert it == 0 and np.array_equal(x, xs)
    x, it, res = dev.pcg(np.zeros(len(b)), b, 100, 1e-10)          # and the ordinary case still works
    assert 0 < it < 15 and np.all(np.isfinite(x)) and res[-1] < 1e-10 * np.linalg.norm(b)
    bad = b.copy()
    bad[3] = np.nan
    with pytest.raises(capi.Amg1dError):
        dev.pcg(np.zeros(len(b)), bad, 10, 1e-10)


def test_failed_level_upload_can_be_retried(lib):
    """A level whose upload fails after its device allocation leaves no allocation behind: the byte
    accounting returns to its old value and the same level can be set again."""
    import ctypes as C
    h = C.c_void_p()
    capi.check(None, lib.amg1d_create(C.byref(h), 1, 0, None))
    n, m = 8, 2
    di = np.tile(np.eye(m).ravel(), n)
    z = np.zeros(n * m * m)
    before = lib.amg1d_get_info(h, b"device_bytes")
    perm = np.arange(n * m, dtype=np.int64)
    perm[5] = 10 ** 6                                   # out of range: rejected after the operator allocation
    rc = lib.amg1d_set_level(h, 0, n, m, capi.dptr(z), capi.dptr(di), capi.dptr(z), capi.dptr(di), 0,
                             capi.iptr(perm), n * m)
    assert rc == capi.ERR_ARG
    assert lib.amg1d_get_info(h, b"device_bytes") == before
    rc = lib.amg1d_set_level(h, 0, n, m, capi.dptr(z), capi.dptr(di), capi.dptr(z), capi.dptr(di), 0, None, n * m)
    assert rc == capi.OK
    capi.check(h, lib.amg1d_finalize(h))
    lib.amg1d_destroy(h)


@pytest.mark.parametrize("shape", ["C2_dg3_agg", "cg_heirarchy"])
def test_pipelined_batch_equals_single_calls(shape):
    """amg1d_vcycle_batch (copy-in / compute / copy-out overlapped over PCIe) gives, problem by problem, the bits
    of amg1d_vcycle and - with zero_guess - of amg1d_ldiv; DG level 0 (plain copies) and CG level 0 (permuted)."""
    Hp, _, b = build_package(**SHAPES[shape])
    dev = Hp.device
    rng = np.random.default_rng(11)
    K = 5
    bs = [b * (1.0 + 0.25 * k) + 1e-3 * rng.standard_normal(len(b)) for k in range(K)]
    x0 = [rng.standard_normal(len(b)) for _ in range(K)]
    ref = [dev.vcycle(x0[k], bs[k], nPre=2, nPost=3, alpha=0.6) for k in range(K)]
    xs = [v.copy() for v in x0]
    dev.vcycle_batch(xs, bs, nPre=2, nPost=3, alpha=0.6)
    for k in range(K):
        assert np.array_equal(xs[k], ref[k]), k
    ref0 = [dev.ldiv(bs[k]) for k in range(K)]
    ys = [np.full(len(b), np.nan) for _ in range(K)]            # never read in the zero-guess form
    dev.vcycle_batch(ys, bs, zero_guess=True)
    for k in range(K):
        assert np.array_equal(ys[k], ref0[k]), k
    dev.vcycle_batch([], [])                                    # an empty batch is fine
    assert np.array_equal(dev.vcycle(x0[0], bs[0], nPre=2, nPost=3, alpha=0.6), ref[0])   # the handle is still sane
    dev.close()
