"""Oracle (test infrastructure): ctypes driver of oracle/vcycle_ref.c, the plain-C restatement of
src/solvers.jl:19-50.  Feeds it an oracle ``MeshHierarchy`` (scipy operators, LU blocks)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from .smoother import JacobiSmoother

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvcycle_ref.so")
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int64)


def load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "vcycle_ref.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    lib = C.CDLL(_LIB)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int]
    lib.ref_set_level.argtypes = [C.c_void_p, C.c_int, C.c_int64, _pi, _pi, _pd, C.c_int, C.c_int64, _pi, _pd, _pd]
    lib.ref_set_transfer.argtypes = [C.c_void_p, C.c_int, C.c_int64, _pi, _pi, _pd, _pi, _pi, _pd]
    lib.ref_set_coarse_dense.argtypes = [C.c_void_p, C.c_int64, _pd]
    lib.ref_vcycle.argtypes = [C.c_void_p, _pd, _pd, C.c_int, C.c_int, C.c_double]
    lib.ref_residual_norm.restype = C.c_double
    lib.ref_residual_norm.argtypes = [C.c_void_p, _pd, _pd]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_set_threads.argtypes = [C.c_int]
    lib.ref_set_threads.restype = C.c_int
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


class CRefHierarchy:
    def __init__(self, H):
        self.lib = load()
        self.keep = []
        nL = len(H.mStiffness)
        self.h = self.lib.ref_create(nL)
        for l in range(nL):
            A = sp.csr_matrix(H.mStiffness[l])
            A.sort_indices()
            Ap, Aj, Ax = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
            S = H.mSmoothers[l]
            if isinstance(S, JacobiSmoother):
                jac = np.ascontiguousarray(S.mJac, dtype=np.float64)
                self.keep += [Ap, Aj, Ax, jac]
                rc = self.lib.ref_set_level(self.h, l, A.shape[0], _p(Ap, _pi), _p(Aj, _pi), _p(Ax, _pd), 0, 0,
                                            None, None, _p(jac, _pd))
            else:
                inds = np.ascontiguousarray(S.mBlockInds.T, dtype=np.int64)          # (nblocks, m)
                nb, m = inds.shape
                Ac = sp.csc_matrix(H.mStiffness[l])
                blocks = np.ascontiguousarray(
                    np.stack([Ac[i, :][:, i].toarray() for i in inds]) if nb <= 4096 else _diag_blocks(A, inds))
                self.keep += [Ap, Aj, Ax, inds, blocks]
                rc = self.lib.ref_set_level(self.h, l, A.shape[0], _p(Ap, _pi), _p(Aj, _pi), _p(Ax, _pd), m, nb,
                                            _p(inds, _pi), _p(blocks, _pd), None)
            assert rc == 0, rc
            if l < nL - 1:
                L = sp.csr_matrix(H.mInterpolation[l]); L.sort_indices()
                LT = sp.csr_matrix(sp.csc_matrix(H.mInterpolation[l]).T); LT.sort_indices()
                arrs = [L.indptr.astype(np.int64), L.indices.astype(np.int64), L.data.astype(np.float64),
                        LT.indptr.astype(np.int64), LT.indices.astype(np.int64), LT.data.astype(np.float64)]
                self.keep += arrs
                self.lib.ref_set_transfer(self.h, l, L.shape[1], _p(arrs[0], _pi), _p(arrs[1], _pi), _p(arrs[2], _pd),
                                          _p(arrs[3], _pi), _p(arrs[4], _pi), _p(arrs[5], _pd))
        Ad = np.ascontiguousarray(sp.csr_matrix(H.mStiffness[-1]).toarray())
        self.keep.append(Ad)
        assert self.lib.ref_set_coarse_dense(self.h, Ad.shape[0], _p(Ad, _pd)) == 0

    def vcycle(self, x0, b, nPre=3, nPost=3, alpha=2.0 / 3.0):
        x = np.array(x0, dtype=np.float64, copy=True)
        b = np.ascontiguousarray(b, dtype=np.float64)
        self.lib.ref_vcycle(self.h, _p(x, _pd), _p(b, _pd), nPre, nPost, alpha)
        return x

    def residual_norm(self, x, b):
        return self.lib.ref_residual_norm(self.h, _p(np.ascontiguousarray(x), _pd), _p(np.ascontiguousarray(b), _pd))

    def set_threads(self, n):
        """OpenMP threads for the row-parallel loops (0 = leave unchanged); returns the count in use."""
        return self.lib.ref_set_threads(int(n))

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None


def _diag_blocks(A, inds):
    """Dense diagonal blocks of a CSR matrix for contiguous element index sets (vectorised)."""
    nb, m = inds.shape
    coo = A.tocoo()
    elem = np.full(A.shape[0], -1, dtype=np.int64)
    loc = np.zeros(A.shape[0], dtype=np.int64)
    elem[inds.ravel()] = np.repeat(np.arange(nb), m)
    loc[inds.ravel()] = np.tile(np.arange(m), nb)
    same = elem[coo.row] == elem[coo.col]
    out = np.zeros((nb, m, m))
    np.add.at(out, (elem[coo.row[same]], loc[coo.row[same]], loc[coo.col[same]]), coo.data[same])
    return out
