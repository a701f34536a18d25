"""Replay of julia/device_hierarchy.jl's upload through the C ABI.

The Julia binding cannot be executed here (no Julia in the image), and signature checks
(tests/test_bindings_consistency.py) do not catch an index-base or memory-layout slip.  This test rebuilds, in
numpy, EXACTLY the arrays that file hands to ccall - 1-based slot maps turned into `vec(slots) .- 1`, blocks as
Julia's column-major `(m, m, n)` arrays, `cat(inv(...); dims = 3)` smoother blocks, `(m, n)` reciprocal
diagonals, 0-based parents from `pmin .- 1`, `P0 (mf, mc, nf)` - feeds their raw memory to amg1d_set_level /
amg1d_set_transfer, and requires the solve to reproduce the tested Python upload bit for bit.  Every helper
below mirrors the Julia function of the same name, statement by statement, in 1-based arithmetic."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import _capi as capi
from agglomerationmultigrid1d_b200.cg_mesh import CgMesh
from agglomerationmultigrid1d_b200.smoother import JacobiSmoother
from shapes import SHAPES, build_package

pytestmark = pytest.mark.gpu


def jl_level_slots(mesh):
    """level_slots: slots[i, e] = 1-based DOF of row i of block e, 0 = padding; Julia (m, n) column-major."""
    nodes = np.asarray(mesh.mNodesInd) + 1                         # the reference's 1-based mNodesInd
    if isinstance(mesh, CgMesh):
        n, p = nodes.shape[0], mesh.mP
        s = np.zeros((p, n + 1), dtype=np.int64, order="F")
        for k in range(1, n + 1):
            s[0, k - 1] = nodes[k - 1, 0]
            s[1:p, k - 1] = nodes[k - 1, 2:]
        s[0, n] = nodes[n - 1, 1]
        return s
    return np.asfortranarray(nodes.T.astype(np.int64))


def jl_level_blocks(A, slots):
    m, n = slots.shape
    elemOf = np.zeros(A.shape[0] + 1, dtype=np.int64)
    locOf = np.zeros(A.shape[0] + 1, dtype=np.int64)
    for e in range(1, n + 1):
        for i in range(1, m + 1):
            if slots[i - 1, e - 1] > 0:
                elemOf[slots[i - 1, e - 1]] = e
                locOf[slots[i - 1, e - 1]] = i
    lo, di, up = (np.zeros((m, m, n), order="F") for _ in range(3))
    A = sp.csc_matrix(A)
    for col in range(1, A.shape[1] + 1):
        for k in range(A.indptr[col - 1], A.indptr[col]):
            r = A.indices[k] + 1
            d = elemOf[col] - elemOf[r]
            assert abs(d) <= 1 or A.data[k] == 0.0
            blk = lo if d == -1 else (di if d == 0 else up)
            blk[locOf[r] - 1, locOf[col] - 1, elemOf[r] - 1] += A.data[k]
    for e in range(1, n + 1):
        for i in range(1, m + 1):
            if slots[i - 1, e - 1] == 0:
                di[i - 1, i - 1, e - 1] = 1.0
    return lo, di, up


def jl_smoother_inverse(S, A, slots):
    m, n = slots.shape
    if isinstance(S, JacobiSmoother):
        d = sp.csc_matrix(A).diagonal()
        out = np.ones((m, n), order="F")
        for e in range(n):
            for i in range(m):
                if slots[i, e] > 0:
                    out[i, e] = 1.0 / d[slots[i, e] - 1]
        return out, 1
    Ad = sp.csc_matrix(A)
    blocks = [np.linalg.inv(Ad[np.ix_(slots[:, e] - 1, slots[:, e] - 1)].toarray()) for e in range(n)]   # inv(Matrix(b))
    return np.asfortranarray(np.stack(blocks, axis=2)), 0


def jl_transfer_blocks(L, fs, cs):
    mf, nf = fs.shape
    mc, nc = cs.shape
    L = sp.csc_matrix(L)
    eF = np.zeros(L.shape[0] + 1, dtype=np.int64); lF = np.zeros(L.shape[0] + 1, dtype=np.int64)
    eC = np.zeros(L.shape[1] + 1, dtype=np.int64); lC = np.zeros(L.shape[1] + 1, dtype=np.int64)
    for e in range(1, nf + 1):
        for i in range(1, mf + 1):
            if fs[i - 1, e - 1] > 0:
                eF[fs[i - 1, e - 1]] = e; lF[fs[i - 1, e - 1]] = i
    for e in range(1, nc + 1):
        for i in range(1, mc + 1):
            if cs[i - 1, e - 1] > 0:
                eC[cs[i - 1, e - 1]] = e; lC[cs[i - 1, e - 1]] = i
    big = np.iinfo(np.int64).max
    pmin = np.full(nf + 1, big, dtype=np.int64); pmax = np.zeros(nf + 1, dtype=np.int64)
    for col in range(1, L.shape[1] + 1):
        for k in range(L.indptr[col - 1], L.indptr[col]):
            if L.data[k] == 0.0:
                continue
            e = eF[L.indices[k] + 1]
            pmin[e] = min(pmin[e], eC[col]); pmax[e] = max(pmax[e], eC[col])
    for e in range(1, nf + 1):
        if pmax[e] == 0:
            pmin[e] = pmin[e - 1] if e > 1 else 1
            pmax[e] = pmin[e]
    assert np.all(pmax[1:] - pmin[1:] <= 1) and np.all(np.diff(pmin[1:]) >= 0)
    two = bool(np.any(pmax[1:] > pmin[1:]))
    P0 = np.zeros((mf, mc, nf), order="F")
    P1 = np.zeros((mf, mc, nf), order="F") if two else None
    for col in range(1, L.shape[1] + 1):
        for k in range(L.indptr[col - 1], L.indptr[col]):
            if L.data[k] == 0.0:
                continue
            r = L.indices[k] + 1
            e = eF[r]
            blk = P0 if eC[col] == pmin[e] else P1
            blk[lF[r] - 1, lC[col] - 1, e - 1] += L.data[k]
    return np.ascontiguousarray(pmin[1:] - 1), P0, P1


def raw(a, typ):
    """Pointer to the array's memory exactly as Julia would pass it (column-major storage)."""
    assert a.flags["F_CONTIGUOUS"] or a.ndim == 1
    return a.ctypes.data_as(C.POINTER(typ))


@pytest.mark.parametrize("shape", ["dg_heirarchy", "full_heirarchy", "C4_cg3_dg1_agg"])
def test_julia_style_upload_reproduces_the_python_upload(shape, lib):
    Hp, _, b = build_package(**SHAPES[shape])
    nL = len(Hp.mMeshes)
    h = C.c_void_p()
    capi.check(None, lib.amg1d_create(C.byref(h), nL, 0, None))
    slots = [jl_level_slots(m) for m in Hp.mMeshes]
    keep = []
    for l in range(nL):
        lo, di, up = jl_level_blocks(Hp.mStiffness[l], slots[l])
        dinv, isdiag = jl_smoother_inverse(Hp.mSmoothers[l], Hp.mStiffness[l], slots[l])
        m, n = slots[l].shape
        perm = np.ascontiguousarray(slots[l].ravel(order="F") - 1)                  # vec(slots) .- 1
        permptr = None if np.array_equal(perm, np.arange(m * n)) else raw(perm, C.c_int64)
        keep += [lo, di, up, dinv, perm]
        capi.check(h, lib.amg1d_set_level(h, l, n, m, raw(lo, C.c_double), raw(di, C.c_double), raw(up, C.c_double),
                                          raw(dinv, C.c_double), isdiag, permptr, Hp.mStiffness[l].shape[0]))
    for l in range(nL - 1):
        parent, P0, P1 = jl_transfer_blocks(Hp.mInterpolation[l], slots[l], slots[l + 1])
        keep += [parent, P0, P1]
        capi.check(h, lib.amg1d_set_transfer(h, l, slots[l].shape[1], slots[l].shape[0], slots[l + 1].shape[0],
                                             raw(parent, C.c_int64), raw(P0, C.c_double),
                                             None if P1 is None else raw(P1, C.c_double)))
    capi.check(h, lib.amg1d_finalize(h))
    x = np.zeros(len(b))
    res = np.zeros(100)
    it = C.c_int(0)
    capi.check(h, lib.amg1d_solve(h, capi.dptr(x), capi.dptr(np.ascontiguousarray(b)), 100, 1e-10, 3, 3, 2.0 / 3.0,
                                  C.byref(it), capi.dptr(res), None, None))
    x_py, it_py, res_py, _ = aggmg.multigrid(Hp, np.zeros(len(b)), b, 100, 1e-10, with_error=False)
    assert it.value == it_py
    assert np.array_equal(res[:it_py], res_py) and np.array_equal(x, x_py)
    lib.amg1d_destroy(h)
    Hp.device.close()
