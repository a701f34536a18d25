"""The level-0 legs of a DG p = 3 hierarchy exist in four forms (csrc/kernels_fused.cuh), selected by options:

  recompute_dinv = 0                      f_down / f_up streaming the stored block-Jacobi inverse
  recompute_dinv = 4, leg_pipeline = 0    f_down / f_up inverting A_di in registers, inverse kept in shared memory
  ... + dinv_registers = 1                f_down_dv / f_up_dv: the inverse stays in registers
  recompute_dinv = 4, leg_pipeline = 1    f_down_pp / f_up_pp: persistent CTAs, next window prefetched with TMA bulk
                                          copies + mbarrier (the default)
  leg_pipeline = 2 / dinv_registers = 2   the same two forms on the 2 x 2 levels below as well (measured slower, kept
                                          as options)

They perform the same arithmetic in the same order, so iterates, residual norms (same window partition of the
two-stage reduction) and whole solves must be BIT-IDENTICAL - on sizes that are not a multiple of the window, with
fewer windows than persistent CTAs and with many windows per CTA, with and without a zero initial guess (ldiv!),
for several sweep counts.  The default form is compared with the oracle in tests/test_gpu_parity.py /
tests/test_gpu_atscale.py."""
import math

import numpy as np
import pytest

from agglomerationmultigrid1d_b200 import uniform

pytestmark = pytest.mark.gpu

SWEEPS = ((3, 3, 2.0 / 3.0), (1, 1, 1.0), (2, 0, 0.5), (0, 2, 0.7), (5, 4, 0.6))
VARIANTS = (
    ("streamed", dict(recompute_dinv=0, leg_pipeline=0, dinv_registers=0)),
    ("recompute_smem", dict(recompute_dinv=4, leg_pipeline=0, dinv_registers=0)),
    ("recompute_regs", dict(recompute_dinv=4, leg_pipeline=0, dinv_registers=1)),
    ("pipelined", dict(recompute_dinv=4, leg_pipeline=1, dinv_registers=0)),
    ("pipelined_2x2_too", dict(recompute_dinv=4, leg_pipeline=2, dinv_registers=0)),
    ("registers_2x2_too", dict(recompute_dinv=4, leg_pipeline=0, dinv_registers=2)),
)


def _set(dev, opts):
    for k, v in opts.items():
        dev.set_option(k, v)


def _run(dev, x0, b):
    out = {("v", a, c): dev.vcycle(x0, b, nPre=a, nPost=c, alpha=al) for a, c, al in SWEEPS}
    out["ldiv"] = dev.ldiv(b)
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 30, 1e-10)
    out["solve_x"], out["solve_it"], out["solve_res"] = x, it, np.asarray(res)
    dev.dev_set_problem(x0, b)
    dev.dev_vcycle(with_residual_norm=True)
    out["dev_norm"] = dev.dev_residual_norm()
    out["launches"] = dev.info("launches_per_cycle")
    return out


# 3 * 2^k and 5 * 2^k elements: not a multiple of the 120-element window; 1536: 13 windows (fewer than persistent
# CTAs); 786432: 6554 windows = 11 per persistent CTA on 148 SMs x 4
@pytest.mark.parametrize("orders,n", [((3, 1), 1536), ((3, 1), 40960), ((3, 1), 786432), ((3, 2, 1), 98304)])
def test_leg_variants_bit_identical(orders, n):
    k = (n & -n).bit_length() - 1
    U = uniform.UniformDgHierarchy(n, list(orders), [2] * k, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    try:
        w = 2.0 * math.pi / 64.0
        b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
        x0 = np.random.default_rng(5).standard_normal(len(b))
        # the default: pipelined from leg_pipeline_min elements up (smaller levels keep the one-window-per-CTA legs)
        assert dev.info("leg_pipeline") == 1
        assert dev.info("leg_pipeline:0") == (1 if n >= dev.info("leg_pipeline_min") else 0)
        dev.set_option("leg_pipeline_min", 0)                      # here: every size through every form
        ref = None
        for name, opts in VARIANTS:
            _set(dev, opts)
            assert dev.info("leg_pipeline:0") == (1 if name.startswith("pipelined") else 0)
            assert dev.info("dinv_recompute:0") == (0 if name == "streamed" else 1)
            got = _run(dev, x0, b)
            if ref is None:
                ref = got
                continue
            for key, val in ref.items():
                if isinstance(val, np.ndarray):
                    assert np.array_equal(got[key], val), (name, key)
                else:
                    assert got[key] == val, (name, key)
        assert ref["solve_it"] < 30
    finally:
        dev.close()


def test_pipelined_legs_on_a_graded_mesh():
    """Explicit per-element blocks (MeshHierarchy: global sparse algebra, no pattern upload): every element has its own
    operator, so a window that took a wrong tile or lane from the staged copy cannot go unnoticed."""
    import agglomerationmultigrid1d_b200 as aggmg
    n = 3 * 2 ** 12
    rng = np.random.default_rng(0)
    x = np.concatenate([[0.0], np.cumsum(0.5 + rng.random(n))])
    xout = float(x[-1])
    w = 2.0 * math.pi / 64.0
    mesh = aggmg.Mesh(x)
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
    H = aggmg.MeshHierarchy(meshes, [bd] * len(meshes), A, G, D, C, nDG=2, nAgg=10, upload=False)
    dev = H.upload()
    try:
        dev.set_option("leg_pipeline_min", 0)
        assert dev.info("pattern:0") == 0 and dev.info("leg_pipeline:0") == 1
        x0 = np.random.default_rng(7).standard_normal(len(b))
        ref = None
        for name, opts in VARIANTS:
            _set(dev, opts)
            got = dev.vcycle(x0, b), dev.ldiv(b)
            if ref is None:
                ref = got
            else:
                assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), name
    finally:
        dev.close()
