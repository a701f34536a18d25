"""Host set-up: Legendre basis, Gauss quadrature and the nodal reference element (vectorised).

Mirrors src/legendre.jl:14-58, src/gauss_quad.jl:6-12, src/reference_element.jl:15-90 of the
reference.  Set-up stays on the host (BASELINE north_star); nothing here is on the hot path.
"""
import math

import numpy as np


def legendre_val_and_deriv(x, n):
    """P_0..P_n and derivatives at the points x; returns two (len(x), n+1) arrays
    (src/legendre.jl:43-58: d_i = (2i-1) P_{i-1} + d_{i-2})."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    f = np.zeros((len(x), n + 1))
    d = np.zeros((len(x), n + 1))
    f[:, 0] = 1.0
    if n >= 1:
        f[:, 1] = x
        d[:, 1] = 1.0
    for i in range(2, n + 1):
        f[:, i] = ((2 * i - 1) * x * f[:, i - 1] - (i - 1) * f[:, i - 2]) / i
        d[:, i] = (2 * i - 1) * f[:, i - 1] + d[:, i - 2]
    return f, d


def legendre_val(x, n):
    return legendre_val_and_deriv(x, n)[0]


def gauss_quad(p):
    """Gauss nodes / weights on [-1, 1] exact to degree p (Golub-Welsch, src/gauss_quad.jl:6-12)."""
    n = int(math.ceil((p + 1) / 2))
    k = np.arange(1, n, dtype=np.float64)
    beta = k / np.sqrt(4.0 * k * k - 1.0)
    ev, evec = np.linalg.eigh(np.diag(beta, 1) + np.diag(beta, -1))
    return ev, 2.0 * evec[0, :] ** 2


def evaluate_nodal_basis_fun_and_deriv(basisFunCoeff, nodes):
    """(len(nodes), p+1) values and derivatives of the nodal basis (src/reference_element.jl:60-90)."""
    p = basisFunCoeff.shape[0] - 1
    leg, dleg = legendre_val_and_deriv(nodes, p)
    return leg @ basisFunCoeff, dleg @ basisFunCoeff


def evaluate_nodal_basis_fun(basisFunCoeff, nodes):
    return evaluate_nodal_basis_fun_and_deriv(basisFunCoeff, nodes)[0]


class ReferenceElement:
    """Nodes -1, +1, cos(pi k / p) (k = 1..p-1); Vandermonde-inverse coefficients; Gauss rule of
    degree 2p; reference mass matrix (src/reference_element.jl:15-54)."""

    def __init__(self, mP):
        self.mP = int(mP)
        if mP >= 1:
            self.mNodesX = np.concatenate(([-1.0, 1.0], np.cos(np.pi * np.arange(1, mP) / mP)))
        else:
            self.mNodesX = np.array([0.0])
        self.mBasisFunCoeff = np.linalg.inv(legendre_val(self.mNodesX, mP))
        self.mGaussQuadNodes, self.mGaussQuadWeights = gauss_quad(2 * mP)
        self.mBasisGQFunVal, self.mBasisGQDerivVal = evaluate_nodal_basis_fun_and_deriv(
            self.mBasisFunCoeff, self.mGaussQuadNodes)
        w = self.mGaussQuadWeights
        M = np.einsum("l,li,lj->ij", w, self.mBasisGQFunVal, self.mBasisGQFunVal)
        self.mMassMatrix = np.triu(M) + np.triu(M, 1).T
