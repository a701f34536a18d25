"""Why the fused legs may invert the block-Jacobi blocks WITHOUT pivoting (csrc/device_setup.cuh: k_dinv_recompute,
csrc/kernels_fused.cuh: reg_invert<M, false>), shown on the CPU with the statements of those kernels in numpy:

* the diagonal blocks of the reference's DG operators (end nodes numbered first, src/dg_mesh.jl:41-46) DO swap rows under
  partial pivoting - in every element - so the device cannot simply assume "no pivots";
* they are symmetric positive definite, and Gauss-Jordan elimination without pivoting gives the same inverse to far below
  the 1e-12 the device demands before a level adopts the unpivoted form (uniform and graded meshes, p = 1 .. 4, and the
  agglomerated levels underneath);
* a block that really needs its pivots (tiny leading entry) fails that comparison, i.e. the level would keep the
  partial-pivoting chain.
The device applies the test element by element at upload (flag 4 of k_dinv_recompute); the GPU side of it -
bit-identical iterates between the pivoted / unpivoted / streamed forms - is tests/test_gpu_leg_variants.py."""
import math

import numpy as np
import pytest

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import blocks as blk

NOPIVOT_TOL = 1e-12          # AMG1D_NOPIVOT_TOL


def gj_pivoted(a):
    """k_dinv_recompute modes 0-2: compare-and-swap chain, elimination, column swaps undone in reverse.  a: (m, m),
    a[r, c]; returns (inverse, swapped)."""
    a = a.copy()
    m = a.shape[0]
    sw = np.zeros((m, m), dtype=bool)
    swapped = False
    for c in range(m):
        for r in range(c + 1, m):
            s = abs(a[r, c]) > abs(a[c, c])
            sw[c, r] = s
            swapped = swapped or s
            if s:
                a[[c, r], :] = a[[r, c], :]
        dd = 1.0 / a[c, c]
        a[c, c] = 1.0
        a[c, :] *= dd
        for r in range(m):
            if r != c:
                f = a[r, c]
                a[r, c] = 0.0
                a[r, :] += -f * a[c, :]
    for c in range(m - 1, -1, -1):
        for r2 in range(m - 1, c, -1):
            if sw[c, r2]:
                a[:, [c, r2]] = a[:, [r2, c]]
    return a, swapped


def gj_unpivoted(a):
    """gj_invert_nopivot / reg_invert<M, false>"""
    a = a.copy()
    m = a.shape[0]
    for c in range(m):
        if a[c, c] == 0.0:
            return None
        dd = 1.0 / a[c, c]
        a[c, c] = 1.0
        a[c, :] *= dd
        for r in range(m):
            if r != c:
                f = a[r, c]
                a[r, c] = 0.0
                a[r, :] += -f * a[c, :]
    return a


def _hierarchy(vertices, p, n_agg):
    n = len(vertices) - 1
    xout = float(vertices[-1])
    w = 2.0 * math.pi / 64.0
    mesh = aggmg.Mesh(np.asarray(vertices, dtype=float))
    bd = aggmg.set_boundary(mesh, 0.0, xout, [("neu", 0.0), ("dir", math.cos(w * xout))])
    orders = [p] + ([1] if p > 1 else [])
    meshes = [aggmg.DgMesh(mesh, q) for q in orders]
    cur = n
    for i in range(n_agg):
        agg = [[2 * j, 2 * j + 1] for j in range(cur // 2)]
        cur //= 2
        meshes.append(aggmg.AgglomeratedDgMesh1(1, agg, mesh, meshes[len(orders) - 1]) if i == 0
                      else aggmg.AgglomeratedDgMeshN(1, agg, meshes[-1], meshes[len(orders) - 1]))
    G, D, C = aggmg.dg_flux_operators(meshes[0], mesh, bd, 1000.0)
    A = (C - D @ meshes[0].mMassMatrixLU.solve(G)).tocsc()
    return aggmg.MeshHierarchy(meshes, [bd] * len(meshes), A, G, D, C, nDG=len(orders), nAgg=n_agg, upload=False)


def _diag_blocks(H, level):
    slots = blk.level_slots(H.mMeshes[level])
    _, di, _ = blk.csc_to_blocks(H.mStiffness[level], slots)
    di = np.asarray(di)
    n = di.shape[0]
    m = int(round(math.sqrt(di.size // n)))
    return di.reshape(n, m, m)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("graded", [False, True])
def test_dg_blocks_swap_rows_but_do_not_need_to(p, graded):
    n = 64
    rng = np.random.default_rng(p)
    x = np.concatenate([[0.0], np.cumsum(0.25 + rng.random(n))]) if graded else np.arange(n + 1, dtype=float)
    H = _hierarchy(x, p, 3)
    any_swap_level0 = False
    for level in range(len(H.mMeshes)):
        worst = 0.0
        for a in _diag_blocks(H, level):
            # the layout convention does not matter for the claim (a block and its transpose pivot alike up to symmetry):
            # check both
            for blk_ in (a, a.T):
                piv, swapped = gj_pivoted(blk_)
                unp = gj_unpivoted(blk_)
                assert unp is not None
                scale = np.abs(piv).max()
                worst = max(worst, np.abs(unp - piv).max() / scale)
                assert np.abs(piv - np.linalg.inv(blk_)).max() <= 1e-12 * scale
                if level == 0:
                    any_swap_level0 = any_swap_level0 or swapped
            assert np.allclose(a, a.T, rtol=1e-12, atol=1e-12 * np.abs(a).max())       # symmetric
            assert np.linalg.eigvalsh(0.5 * (a + a.T)).min() > 0.0                     # positive definite
        assert worst <= 0.01 * NOPIVOT_TOL, (level, worst)      # two orders of margin below the device's bar
    if p >= 2:
        assert any_swap_level0      # end-nodes-first numbering: the largest entry of a column is not on the diagonal


def test_a_block_that_needs_its_pivots_fails_the_comparison():
    a = np.array([[1e-18, 1.0, 0.3], [1.0, 0.7, 0.2], [0.3, 0.2, 1.1]])     # 0.7 - 1e18 loses the 0.7 without a row swap
    piv, swapped = gj_pivoted(a)
    unp = gj_unpivoted(a)
    assert swapped
    assert np.abs(piv - np.linalg.inv(a)).max() <= 1e-12 * np.abs(piv).max()
    assert unp is None or not (np.abs(unp - piv).max() <= NOPIVOT_TOL * np.abs(piv).max())
    z = np.array([[0.0, 1.0], [1.0, 0.0]])
    assert gj_unpivoted(z) is None and gj_pivoted(z)[1]
