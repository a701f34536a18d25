"""Oracle (test infrastructure): Legendre values, Gauss quadrature and the nodal reference element.

Follows src/legendre.jl:14-58, src/gauss_quad.jl:6-12, src/reference_element.jl:15-90.
"""
import math

import numpy as np


def legendre_val(x, n):
    """All Legendre polynomials up to degree n at x (src/legendre.jl:14-25)."""
    if n == 0:
        return np.array([1.0])
    f = [1.0, float(x)]
    for i in range(2, n + 1):
        f.append(((2 * i - 1) * x * f[i - 1] - (i - 1) * f[i - 2]) / i)
    return np.array(f)


def legendre_val_and_deriv(x, n):
    """Values and derivatives (src/legendre.jl:43-58); d_i = (2i-1) f_{i-1} + d_{i-2}."""
    if n == 0:
        return np.array([1.0]), np.array([0.0])
    f = [1.0, float(x)]
    d = [0.0, 1.0]
    for i in range(2, n + 1):
        f.append(((2 * i - 1) * x * f[i - 1] - (i - 1) * f[i - 2]) / i)
        d.append((2 * i - 1) * f[i - 1] + d[i - 2])
    return np.array(f), np.array(d)


def gauss_quad(p):
    """Golub-Welsch Gauss quadrature on [-1,1] for degree of precision p (src/gauss_quad.jl:6-12)."""
    n = int(math.ceil((p + 1) / 2))
    b = np.arange(1, n, dtype=np.float64)
    b = b / np.sqrt(4.0 * b * b - 1.0)
    T = np.diag(b, 1) + np.diag(b, -1)
    ev, evec = np.linalg.eigh(T)
    return ev, 2.0 * evec[0, :] ** 2


def evaluate_nodal_basis_fun(coeff, nodes):
    """basisFunVal[l, i] = dot(coeff[:, i], legendre(nodes[l])) (src/reference_element.jl:60-73)."""
    nodes = np.atleast_1d(np.asarray(nodes, dtype=np.float64))
    p = coeff.shape[0] - 1
    val = np.zeros((len(nodes), p + 1))
    for l, x in enumerate(nodes):
        leg = legendre_val(x, p)
        for i in range(p + 1):
            val[l, i] = np.dot(coeff[:, i], leg)
    return val


def evaluate_nodal_basis_fun_and_deriv(coeff, nodes):
    """src/reference_element.jl:75-90."""
    nodes = np.atleast_1d(np.asarray(nodes, dtype=np.float64))
    p = coeff.shape[0] - 1
    val = np.zeros((len(nodes), p + 1))
    der = np.zeros((len(nodes), p + 1))
    for l, x in enumerate(nodes):
        leg, dleg = legendre_val_and_deriv(x, p)
        for i in range(p + 1):
            val[l, i] = np.dot(coeff[:, i], leg)
            der[l, i] = np.dot(coeff[:, i], dleg)
    return val, der


class ReferenceElement:
    """src/reference_element.jl:1-54.  Node order: -1, +1, then cos(pi k / p), k = 1..p-1."""

    def __init__(self, mP):
        self.mP = mP
        x = np.zeros(mP + 1)
        if mP >= 1:
            x[0:2] = [-1.0, 1.0]
            x[2:] = np.cos(np.pi * np.arange(1, mP) / mP)
        else:
            x[0] = 0.0
        self.mNodesX = x
        V = np.zeros((mP + 1, mP + 1))
        for i in range(mP + 1):
            V[i, :] = legendre_val(x[i], mP)
        self.mBasisFunCoeff = np.linalg.inv(V)
        self.mGaussQuadNodes, self.mGaussQuadWeights = gauss_quad(2 * mP)
        self.mBasisGQFunVal, self.mBasisGQDerivVal = evaluate_nodal_basis_fun_and_deriv(
            self.mBasisFunCoeff, self.mGaussQuadNodes)
        M = np.zeros((mP + 1, mP + 1))
        for j in range(mP + 1):
            for i in range(j + 1):
                for l in range(len(self.mGaussQuadWeights)):
                    M[i, j] += (self.mGaussQuadWeights[l] * self.mBasisGQFunVal[l, i]
                                * self.mBasisGQFunVal[l, j])
        for j in range(mP + 1):
            for i in range(j + 1, mP + 1):
                M[i, j] = M[j, i]
        self.mMassMatrix = M
