"""Oracle (test infrastructure): continuous-Galerkin discretisation of -u'' = f.

Follows src/cg_mesh.jl:26-48 (element), :54-80 (mesh), :87-122 (``cg_stiffness``),
:125-185 (``cg_stiffness_and_rhs``), :188-247 (``cg_rhs``).
DOF numbering (0-based here): vertices 0..n, then element k's interior nodes consecutively.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .refelem import ReferenceElement
from .refmesh import isBoundary


class CgElement:
    def __init__(self, face, mP, vertCounter, refEl):
        self.mIndex = face.mIndex
        self.mP = mP
        h = face.mVertices[1].mX - face.mVertices[0].mX
        xc = (face.mVertices[0].mX + face.mVertices[1].mX) / 2.0
        self.mJacobian = h / 2.0
        self.mRefMap = lambda xi, xc=xc, h=h: xc + h / 2.0 * xi
        self.mNodesInd = np.zeros(mP + 1, dtype=np.int64)
        self.mNodesX = np.zeros(mP + 1)
        for i in range(2):
            self.mNodesInd[i] = face.mVertices[i].mIndex - 1
            self.mNodesX[i] = face.mVertices[i].mX
        for i in range(2, mP + 1):
            self.mNodesInd[i] = vertCounter
            self.mNodesX[i] = self.mRefMap(refEl.mNodesX[i])
            vertCounter += 1


class CgMesh:
    def __init__(self, mesh, mP):
        self.mP = mP
        self.mRefEl = ReferenceElement(mP)
        self.mElements = []
        vertCounter = len(mesh.mVertices)          # 0-based id of the first interior node
        for face in mesh.mFaces:
            self.mElements.append(CgElement(face, mP, vertCounter, self.mRefEl))
            vertCounter += mP - 1
        self.mNumNodes = vertCounter
        rows, cols, vals = [], [], []
        for el in self.mElements:
            for j, node2 in enumerate(el.mNodesInd):
                for i, node1 in enumerate(el.mNodesInd):
                    rows.append(node1); cols.append(node2)
                    vals.append(el.mJacobian * self.mRefEl.mMassMatrix[i, j])
        self.mMassMatrix = sp.csc_matrix((vals, (rows, cols)),
                                         shape=(self.mNumNodes, self.mNumNodes))
        self._lu = None

    @property
    def mMassMatrixLU(self):
        if self._lu is None:
            self._lu = spla.splu(self.mMassMatrix)
        return self._lu


def _local_stiffness(el, refEl):
    n = len(el.mNodesInd)
    temp = np.zeros((n, n))
    for j in range(n):
        for i in range(n):
            for l in range(len(refEl.mGaussQuadNodes)):
                temp[i, j] += ((1.0 / el.mJacobian) * refEl.mGaussQuadWeights[l]
                               * refEl.mBasisGQDerivVal[l, i] * refEl.mBasisGQDerivVal[l, j])
    return temp


def _assemble_stiffness(cgMesh):
    refEl = cgMesh.mRefEl
    rows, cols, vals = [], [], []
    for el in cgMesh.mElements:
        temp = _local_stiffness(el, refEl)
        for j, node2 in enumerate(el.mNodesInd):
            for i, node1 in enumerate(el.mNodesInd):
                rows.append(node1); cols.append(node2); vals.append(temp[i, j])
    return sp.csc_matrix((vals, (rows, cols)), shape=(cgMesh.mNumNodes, cgMesh.mNumNodes))


def _strong_dirichlet(A, dirNodes0):
    """Rows and columns of the Dirichlet DOFs zeroed, identity on their diagonal
    (src/cg_mesh.jl:116-119, :179-182)."""
    A = A.tolil()
    for d in dirNodes0:
        A[d, :] = 0.0
        A[:, d] = 0.0
    for d in dirNodes0:
        A[d, d] = 1.0
    A = A.tocsc()
    A.eliminate_zeros()
    return A


def cg_stiffness(cgMesh, bdCond):
    A = _assemble_stiffness(cgMesh)
    return _strong_dirichlet(A, [d - 1 for d in bdCond.mDirNodes])


def _neumann_rhs(f, mesh, bdCond):
    for k in bdCond.mNeuNodes:
        vert = mesh.mVertices[k - 1]
        face = mesh.mFaces[vert.mFaces[0] - 1]
        if vert is face.mVertices[0]:
            f[vert.mIndex - 1] += -bdCond.mBdCond[-vert.mFaces[1] - 1][1]
        else:
            f[vert.mIndex - 1] += bdCond.mBdCond[-vert.mFaces[1] - 1][1]


def cg_stiffness_and_rhs(cgMesh, mesh, func, bdCond):
    refEl = cgMesh.mRefEl
    A = _assemble_stiffness(cgMesh)
    f = np.zeros(cgMesh.mNumNodes)
    for el in cgMesh.mElements:
        for i, node in enumerate(el.mNodesInd):
            for l in range(len(refEl.mGaussQuadNodes)):
                f[node] += (el.mJacobian * refEl.mGaussQuadWeights[l] * refEl.mBasisGQFunVal[l, i]
                            * func(el.mRefMap(refEl.mGaussQuadNodes[l])))
    _neumann_rhs(f, mesh, bdCond)
    dir0 = [d - 1 for d in bdCond.mDirNodes]
    if dir0:
        f += -(A[:, dir0] @ np.asarray(bdCond.mDirVals, dtype=np.float64))
        f[dir0] = bdCond.mDirVals
    A = _strong_dirichlet(A, dir0)
    return A, f


def cg_rhs(cgMesh, mesh, func, bdCond):
    refEl = cgMesh.mRefEl
    f = np.zeros(cgMesh.mNumNodes)
    for k, el in enumerate(cgMesh.mElements):
        for i, vert in enumerate(el.mNodesInd):
            for l in range(len(refEl.mGaussQuadNodes)):
                f[vert] += (el.mJacobian * refEl.mGaussQuadWeights[l] * refEl.mBasisGQFunVal[l, i]
                            * func(el.mRefMap(refEl.mGaussQuadNodes[l])))
        face = mesh.mFaces[k]
        if isBoundary(face):
            for i, vert1 in enumerate(face.mVertices):
                if isBoundary(vert1) and bdCond.mBdCond[-vert1.mFaces[1] - 1][0] == "dir":
                    for j, node2 in enumerate(el.mNodesInd):
                        temp = 0.0
                        for l in range(len(refEl.mGaussQuadNodes)):
                            temp += ((1.0 / el.mJacobian) * refEl.mGaussQuadWeights[l]
                                     * refEl.mBasisGQDerivVal[l, i] * refEl.mBasisGQDerivVal[l, j])
                        f[node2] += -temp * bdCond.mBdCond[-vert1.mFaces[1] - 1][1]
    _neumann_rhs(f, mesh, bdCond)
    f[[d - 1 for d in bdCond.mDirNodes]] = bdCond.mDirVals
    return f
