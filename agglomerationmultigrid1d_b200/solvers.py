"""Solver entry points: host mirror of src/solvers.jl and of ``apply_smoother`` (src/smoother.jl).

Same names, argument meaning and return values as the reference; every one of them is a call into
libamg1d.so (the GPU).  There is no CPU implementation behind them.

    multigrid_v_cycle(H, x0, b; nPre=3, nPost=3, alpha=2/3) -> x            src/solvers.jl:19-50
    ldiv(H, b) -> b overwritten / ldiv(y, H, b) -> y overwritten            src/solvers.jl:63-92
    pcg(H, x0, b, maxiter, tol) -> (x, iter, res)      CG with ldiv! as preconditioner (the hook's purpose)
    multigrid(H, x0, b, maxiter, tol) -> (x, iter, res, err)                src/solvers.jl:116-139
    iterative_smoother_solve(A, smoother, x0, b; maxiter, tol, alpha)       src/solvers.jl:189-213
    apply_smoother(S, B; alpha) -> alpha * S^-1 B                           src/smoother.jl:52-81
"""
import numpy as np
import scipy.sparse as sp

from . import blocks as blk
from .device import DeviceHierarchy
from .smoother import AbstractSmoother, AdditiveSchwarzSmoother, schwarz_tridiag_blocks, smoother_inverse


def _device_of(H):
    if H.device is None:
        raise RuntimeError("MeshHierarchy was built with upload=False; call H.upload() first")
    return H.device


def multigrid_v_cycle(H, x0, b, nPre=3, nPost=3, alpha=2.0 / 3.0):
    return _device_of(H).vcycle(x0, b, nPre=nPre, nPost=nPost, alpha=alpha)


def ldiv(*args):
    """``ldiv!(H, b)``: b <- one V-cycle from a zero guess; ``ldiv!(y, H, b)``: y <- the same."""
    if len(args) == 2:
        H, b = args
        out = b
    elif len(args) == 3:
        out, H, b = args
    else:
        raise TypeError("ldiv(H, b) or ldiv(y, H, b)")
    out[:] = _device_of(H).ldiv(b)          # zero guess inside the library: no x0 upload, first sweep skips A
    return None


def pcg(H, x0, b, maxiter, tol, nPre=3, nPost=3, alpha=2.0 / 3.0):
    """Conjugate gradients preconditioned with ``ldiv!(z, H, r)`` - the use the reference's ``ldiv!``
    methods are written for (src/solvers.jl:63-92: MeshHierarchy as the ``Pl`` of a Krylov solver); the
    reference itself ships no driver.  Returns (x, iter, res) with res[i] = ||r_i||_2 and the stop rule of
    ``multigrid`` (res < tol ||b||)."""
    return _device_of(H).pcg(x0, b, maxiter, tol, nPre=nPre, nPost=nPost, alpha=alpha)


def multigrid(H, x0, b, maxiter, tol, u_exact=None, with_error=True):
    """Always cycles with nPre = nPost = 3, alpha = 2/3 like the reference (src/solvers.jl:125).
    ``err`` needs u_exact = A \\ b (GPU direct solve, ~10 n m^2 doubles of factors); pass
    ``with_error=False`` to skip it (err is then filled with NaN)."""
    b = np.asarray(b, dtype=np.float64)
    if maxiter <= 0:
        return np.zeros(len(x0)), 0, np.zeros(0), np.zeros(0)
    if u_exact is None and with_error:
        # src/solvers.jl:120: u_exact = A \\ b - block cyclic reduction on the GPU (amg1d_direct_solve)
        u_exact = _device_of(H).direct_solve(0, b)
    return _device_of(H).solve(x0, b, maxiter, tol, u_exact=u_exact)


def _single_level(A, smoother):
    """One-level device hierarchy for a stand-alone (operator, smoother) pair, cached on the smoother."""
    if smoother._device is not None and smoother._device_A is A:
        return smoother._device
    slots = smoother._slots
    if slots is None:
        raise ValueError("smoother was not built by cg_smoother / dg_smoother")
    lo, di, up = blk.csc_to_blocks(A, slots)
    dinv, is_diag = smoother_inverse(smoother, slots)
    dev = DeviceHierarchy(1)
    dev.set_level_blocks(0, lo, di, up, dinv, is_diag, slots, A.shape[0])
    if isinstance(smoother, AdditiveSchwarzSmoother):       # also HybridSchwarzSmoother
        dev.set_level_smoother(0, *schwarz_tridiag_blocks(smoother, slots))
    dev.finalize()
    smoother._device = dev
    smoother._device_A = A
    return dev


def apply_smoother(A, B, alpha=1.0):
    """``apply_smoother(A::AbstractSmoother, B; alpha)``; B may be a vector, a dense matrix or a
    sparse matrix (the scripts pass the operator itself, tests/dg_smoother_test.jl:105)."""
    if not isinstance(A, AbstractSmoother):
        raise TypeError("apply_smoother needs a smoother")
    if sp.issparse(B):
        B = B.toarray()
    if A._owner is not None:
        dev, level = A._owner
        return dev.apply_smoother(level, B, alpha)
    if A._A is None:
        raise ValueError("smoother was not built by cg_smoother / dg_smoother")
    return _single_level(A._A, A).apply_smoother(0, B, alpha)


def iterative_smoother_solve(A, smoother, x0, b, maxiter=1000, tol=1e-6, alpha=1.0, u_exact=None,
                             with_error=True):
    b = np.asarray(b, dtype=np.float64)
    if maxiter <= 0:
        return np.zeros(len(x0)), 0, np.zeros(0), np.zeros(0)
    dev = _single_level(A, smoother)
    if u_exact is None and with_error:
        u_exact = dev.direct_solve(0, b)          # src/solvers.jl:194, on the GPU
    return dev.smoother_solve(0, x0, b, maxiter=maxiter, tol=tol, alpha=alpha, u_exact=u_exact)
