"""The oracle's block-pattern storage (oracle/vcycle_ref.c, oracle/cref.CRefPattern) is the same restatement
as its CSR storage: fed with the same numbers it must reproduce the CSR run BIT FOR BIT - including CG
levels, whose rows are walked in the reference's vertex-first DOF order although the vectors are held in
group order - and it must agree with the independent literal oracle (own assembly, Python V-cycle) to
rounding.  Only then may it judge the GPU at BASELINE's own sizes (tests/test_gpu_atscale.py)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from agglomerationmultigrid1d_b200 import blocks as blk, uniform
from oracle import cref, drivers, solvers
from oracle.hierarchy import MeshHierarchy
from oracle.smoother import BlockJacobi, JacobiSmoother

W = 2.0 * math.pi / 64.0


def _slots(U, l):
    lv = U.levels[l]
    if getattr(lv, "is_cg", False):
        return U.group_slots(l)
    return np.arange(lv.n * lv.m, dtype=np.int64).reshape(lv.n, lv.m)


def _n_host(U, l):
    lv = U.levels[l]
    if getattr(lv, "is_cg", False):
        p = U.cg_orders[l]
        return (lv.n - 1) * p + 1          # n elements: n + 1 vertices + n (p - 1) interior nodes
    return lv.n * lv.m


def _transfer_blocks(U, l):
    if hasattr(U, "transfer_blocks"):
        return U.transfer_blocks(l)
    P, ratio = U.transfers[l]
    e = np.arange(U.levels[l].n)
    return e // ratio, P[e % ratio], None


def csr_hierarchy(U):
    """The hierarchy of U as global sparse matrices in the REFERENCE's DOF numbering (CG: vertex-first)."""
    S, Sm, I = [], [], []
    slots = [_slots(U, l) for l in range(len(U.levels))]
    nh = [_n_host(U, l) for l in range(len(U.levels))]
    for l, lv in enumerate(U.levels):
        lo, di, up = U.level_blocks(l)
        A = blk.blocks_to_csc(lo, di, up, slots[l], nh[l])
        S.append(A)
        if getattr(lv, "is_cg", False):
            Sm.append(JacobiSmoother(A.diagonal()))
        else:
            Sm.append(BlockJacobi(None, slots[l].T))
    for l in range(len(U.levels) - 1):
        parent, P0, P1 = _transfer_blocks(U, l)
        nf, mf, mc = P0.shape
        nc = U.levels[l + 1].n
        rows, cols, vals = [], [], []
        for P, off in ((P0, 0), (P1, 1)):
            if P is None:
                continue
            par = parent + off
            e = np.flatnonzero((par >= 0) & (par < nc))
            r = np.repeat(slots[l][e][:, :, None], mc, axis=2)
            c = np.repeat(slots[l + 1][par[e]][:, None, :], mf, axis=1)
            keep = (r >= 0) & (c >= 0)
            rows.append(r[keep]); cols.append(c[keep]); vals.append(P[e][keep])
        L = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nh[l], nh[l + 1]))
        L.eliminate_zeros()
        I.append(L)
    return MeshHierarchy([None] * len(S), S, None, None, None, Sm, I, None), slots, nh


def to_group(v_host, slots):
    out = np.zeros(slots.size)
    ok = slots.ravel() >= 0
    out[ok] = v_host[slots.ravel()[ok]]
    return out


CASES = {
    "C2_T_C5_dg3": dict(n=256, dg=[3, 1]),
    "C3_dg4": dict(n=256, dg=[4, 2, 1]),
    "P8_dg8": dict(n=128, dg=[8, 4, 2, 1]),
    "C4_cg3_dg1": dict(n=256, cg=[3, 1], dg=[1]),
    "full_cg8": dict(n=256, cg=[8, 4, 2, 1], dg=[], agg=[4] + [2] * 6),
    "C1_cg1": dict(n=256, cg=[1], dg=[], agg=[2]),
}


def build_uniform(n, dg=(), cg=(), agg=None, unit_h=True):
    k = n.bit_length() - 1
    agg = [2] * k if agg is None else agg
    kw = dict(xin=0.0, xout=float(n), CDir=1000.0) if unit_h else dict(xin=0.0, xout=1.0, CDir=1000.0 * n)
    if cg:
        return uniform.UniformCgHierarchy(n, list(cg), list(dg), agg, **kw)
    return uniform.UniformDgHierarchy(n, list(dg), agg, **kw)


@pytest.mark.parametrize("name", sorted(CASES))
def test_pattern_storage_reproduces_csr_storage_bit_for_bit(name):
    kw = CASES[name]
    U = build_uniform(**kw)
    H, slots, nh = csr_hierarchy(U)
    c_csr = cref.CRefHierarchy(H)
    c_pat = cref.CRefPattern(*cref.pattern_arrays(U))
    rng = np.random.default_rng(3)
    b_host = rng.standard_normal(nh[0])
    x_host = rng.standard_normal(nh[0])
    b_grp, x_grp = to_group(b_host, slots[0]), to_group(x_host, slots[0])
    ok = slots[0].ravel() >= 0
    for nPre, nPost, alpha in ((3, 3, 2.0 / 3.0), (1, 2, 0.5), (0, 3, 0.8)):
        y_csr = c_csr.vcycle(x_host, b_host, nPre, nPost, alpha)
        y_pat = c_pat.vcycle(x_grp, b_grp, nPre, nPost, alpha)
        assert np.array_equal(y_pat[ok], y_csr[slots[0].ravel()[ok]]), (name, nPre, nPost)
        assert np.all(y_pat[~ok] == 0.0)
    # the residual vectors are identical; the norm sums them in storage order (group order for CG levels)
    rn_pat, rn_csr = c_pat.residual_norm(x_grp, b_grp), c_csr.residual_norm(x_host, b_host)
    assert abs(rn_pat - rn_csr) <= 1e-14 * rn_csr and (kw.get("cg") or rn_pat == rn_csr)
    c_pat.set_threads(1)                                   # the row walks do not depend on the thread count
    y1 = c_pat.vcycle(x_grp, b_grp)
    c_pat.set_threads(4)
    assert np.array_equal(c_pat.vcycle(x_grp, b_grp), y1)
    c_csr.close(); c_pat.close()


@pytest.mark.parametrize("shape", ["dg3", "cg3"])
def test_pattern_oracle_agrees_with_the_literal_oracle(shape):
    """Independent assembly (oracle/dg.py, cg.py, aggdg.py element loops) + Python V-cycle with SuperLU /
    LAPACK solves against the product-assembled pattern arrays in the C restatement: same V-cycle counts,
    residual histories to 1e-9 (two assemblies, cond(A_coarse) up to 1e8), solutions to 1e-8."""
    n = 256
    if shape == "dg3":
        H, x0, bo, _ = drivers.dg_agg_problem(n, p=3, unit_h=True)
        U = build_uniform(n, dg=[3, 1])
        b = U.rhs(lambda x: W * W * np.cos(W * x), [0.0, math.cos(W * n)])
        slots = np.arange(len(b)).reshape(-1, 4)
    else:
        H, x0, bo, _ = drivers.build_problem(n, cg_orders=[3, 1], dg_orders=[1], agg_factors=[2] * 8)
        U = build_uniform(n, cg=[3, 1], dg=[1], unit_h=False)
        b = U.rhs(np.cos, [-math.sin(0.0), math.cos(1.0)])
        slots = U.group_slots(0)
    ok = slots.ravel() >= 0
    assert np.abs(b[ok] - bo[slots.ravel()[ok]]).max() <= 1e-12 * np.abs(bo).max()
    x_or, it_or, res_or, _ = solvers.multigrid(H, x0, bo, 100, 1e-10)
    c = cref.CRefPattern(*cref.pattern_arrays(U))
    x, it, res = c.multigrid(np.zeros(len(b)), b, 100, 1e-10)
    assert it == it_or
    floor = 1e-13 * np.linalg.norm(bo)
    assert np.all(np.abs(res - res_or) <= np.maximum(1e-9 * res_or, floor)), (res, res_or)
    assert np.abs(x[ok] - x_or[slots.ravel()[ok]]).max() <= 1e-8 * np.abs(x_or).max()
    c.close()
