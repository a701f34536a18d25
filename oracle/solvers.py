"""Oracle (test infrastructure): THE HOT PATH - V-cycle, outer multigrid loop, smoother solve.

Follows src/solvers.jl:19-50 (``multigrid_v_cycle``), :63-92 (``ldiv!`` x2), :116-139
(``multigrid``), :189-213 (``iterative_smoother_solve``).  Every product is a scipy CSC SpMV (the
reference's ``SparseMatrixCSC * Vector``), restriction is ``L' * r``, the coarsest level and the
error reference are sparse direct solves (the reference's ``\\`` = SuiteSparse; here SuperLU).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .smoother import apply_smoother


def _direct_solve(A, b):
    A = sp.csc_matrix(A)
    if A.shape[0] == 1:
        return np.asarray(b, dtype=np.float64) / A[0, 0]
    return spla.spsolve(A, b)


def multigrid_v_cycle(H, x0, b, nPre=3, nPost=3, alpha=2.0 / 3.0):
    n = len(H.mMeshes)
    u = [None] * n
    rhs = [None] * n
    u[0] = np.array(x0, dtype=np.float64)
    rhs[0] = np.asarray(b, dtype=np.float64)
    for k in range(n - 1):
        if k > 0:
            u[k] = np.zeros(H.mStiffness[k].shape[1])
        for _ in range(nPre):
            u[k] = u[k] + apply_smoother(H.mSmoothers[k], rhs[k] - H.mStiffness[k] @ u[k],
                                         alpha=alpha)
        rhs[k + 1] = H.mInterpolation[k].T @ (rhs[k] - H.mStiffness[k] @ u[k])
    u[n - 1] = _direct_solve(H.mStiffness[n - 1], rhs[n - 1])
    for k in range(n - 2, -1, -1):
        u[k] = u[k] + H.mInterpolation[k] @ u[k + 1]
        for _ in range(nPost):
            u[k] = u[k] + apply_smoother(H.mSmoothers[k], rhs[k] - H.mStiffness[k] @ u[k],
                                         alpha=alpha)
    return u[0]


def ldiv(H, b):
    """``ldiv!(H, b)`` / ``ldiv!(y, H, b)``: one V-cycle from a zero initial guess."""
    return multigrid_v_cycle(H, np.zeros(H.mStiffness[0].shape[0]), b)


def pcg(H, x0, b, maxiter, tol, nPre=3, nPost=3, alpha=2.0 / 3.0):
    """Conjugate gradients with M^-1 r = ldiv!(z, H, r) (one V-cycle from zero, src/solvers.jl:84-92).
    The reference only provides the ldiv! hook; the Krylov driver a user would pair it with is
    IterativeSolvers.cg (third-party, not vendored, version unpinned) - this is the textbook
    preconditioned CG it implements (Hestenes-Stiefel recurrences, residual by recurrence).
    Returns (x, iter, res), res[i] = ||r_i||_2; stop rule of ``multigrid``: res < tol ||b||."""
    A = H.mStiffness[0]
    b = np.asarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    r = b - A @ x
    nb = np.linalg.norm(b, 2)
    res = []
    p = None
    rz = 0.0
    for i in range(maxiter):
        z = multigrid_v_cycle(H, np.zeros(len(b)), r, nPre=nPre, nPost=nPost, alpha=alpha)
        rz_new = float(r @ z)
        p = z.copy() if i == 0 else z + (rz_new / rz) * p
        rz = rz_new
        Ap = A @ p
        a = rz / float(p @ Ap)
        x = x + a * p
        r = r - a * Ap
        res.append(np.linalg.norm(r, 2))
        if not res[-1] >= tol * nb:
            break
    return x, len(res), np.array(res)


def multigrid(H, x0, b, maxiter, tol, u_exact=None):
    """Returns (x, iter, res, err).  ``multigrid`` always uses the V-cycle defaults
    nPre = nPost = 3, alpha = 2/3 (src/solvers.jl:125)."""
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros(len(x0))
    if u_exact is None:
        u_exact = _direct_solve(H.mStiffness[0], b)
    err, res = [], []
    nb = np.linalg.norm(b, 2)
    for _ in range(maxiter):
        x = multigrid_v_cycle(H, x0, b)
        x0 = x
        err.append(np.linalg.norm(x - u_exact, 2))
        res.append(np.linalg.norm(H.mStiffness[0] @ x - b, 2))
        if res[-1] < tol * nb:
            break
    return x, len(res), np.array(res), np.array(err)


def iterative_smoother_solve(A, smoother, x0, b, maxiter=1000, tol=1e-6, alpha=1.0, u_exact=None):
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros(len(x0))
    if u_exact is None:
        u_exact = _direct_solve(A, b)
    err, res = [], []
    nb = np.linalg.norm(b, 2)
    x0 = np.asarray(x0, dtype=np.float64)
    for _ in range(maxiter):
        x = x0 + apply_smoother(smoother, b - A @ x0, alpha=alpha)
        x0 = x
        err.append(np.linalg.norm(x - u_exact, 2))
        res.append(np.linalg.norm(A @ x - b, 2))
        if res[-1] < tol * nb:
            break
    return x, len(res), np.array(res), np.array(err)
