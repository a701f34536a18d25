"""Oracle (test infrastructure): nodal discontinuous-Galerkin (LDG-type flux) discretisation.

Follows src/dg_mesh.jl:32-52 (element), :58-138 (mesh, including the switch computation that
compares a face with itself, :96-97, so mSwitch is 1 at every interior vertex),
:144-336 (``dg_flux_operators`` -> G, D, C), :342-457 (``dg_flux_rhs`` -> f, r).
Scripts compose A = C - D (M \\ G), b = f - D (M \\ r) (tests/dg_heirarchy_test.jl:38-42).
DOF numbering (0-based): element k (0-based) owns k(p+1) .. k(p+1)+p; local order = left end,
right end, interior Chebyshev-Lobatto points in descending x.
"""
import numpy as np
import scipy.sparse as sp

from .block_diagonal import BlockDiagonal
from .refelem import ReferenceElement
from .refmesh import isBoundary


class DgElement:
    def __init__(self, face, mP, refEl):
        self.mIndex = face.mIndex
        self.mP = mP
        h = face.mVertices[1].mX - face.mVertices[0].mX
        xc = (face.mVertices[0].mX + face.mVertices[1].mX) / 2.0
        self.mJacobian = h / 2.0
        self.mRefMap = lambda xi, xc=xc, h=h: xc + h / 2.0 * xi
        self.mNodesInd = np.zeros(mP + 1, dtype=np.int64)
        self.mNodesX = np.zeros(mP + 1)
        for i in range(mP + 1):
            self.mNodesInd[i] = (self.mIndex - 1) * (mP + 1) + i
            self.mNodesX[i] = self.mRefMap(refEl.mNodesX[i])


class DgMesh:
    def __init__(self, mesh, mP):
        self.mP = mP
        self.mRefEl = ReferenceElement(mP)
        self.mElements = [DgElement(face, mP, self.mRefEl) for face in mesh.mFaces]
        self.mNumNodes = len(self.mElements) * (mP + 1)
        blocks = [None] * len(self.mElements)
        inds = np.zeros((mP + 1, len(self.mElements)), dtype=np.int64)
        for el in self.mElements:
            blocks[el.mIndex - 1] = el.mJacobian * self.mRefEl.mMassMatrix
            inds[:, el.mIndex - 1] = el.mNodesInd
        self.mMassMatrix = BlockDiagonal(blocks, mP + 1, inds)
        self.mMassMatrixLU = self.mMassMatrix.lu()
        self.mSwitch = compute_switch(mesh.mVertices,
                                      lambda v: mesh.mFaces[v.mFaces[0] - 1].mVertices)


def compute_switch(vertices, first_face_vertices):
    """src/dg_mesh.jl:83-107 (and agglomerated_dg_mesh.jl:360-389, :457-486): 1-based switch."""
    sw = []
    for vert in vertices:
        fv = first_face_vertices(vert)
        if isBoundary(vert):
            sw.append(1 if vert.mX > min(fv[0].mX, fv[1].mX) else 2)
        else:
            # face1 and face2 are BOTH vert.mFaces[1] in the reference, so x1 == x2 -> 1
            x1 = max(fv[0].mX, fv[1].mX)
            x2 = max(fv[0].mX, fv[1].mX)
            sw.append(2 if x1 > x2 else 1)
    return sw


def dg_flux_operators(dgMesh, mesh, bdCond, CDir):
    refEl = dgMesh.mRefEl
    dG, dD, dC = [], [], []
    e1 = 0
    e2 = 1 if dgMesh.mP >= 1 else 0        # p = 0: the single node stands in for both ends
    if dgMesh.mP >= 1:
        for el in dgMesh.mElements:
            n = len(el.mNodesInd)
            temp = np.zeros((n, n))
            for j in range(n):
                for i in range(n):
                    for l in range(len(refEl.mGaussQuadNodes)):
                        temp[i, j] += (refEl.mGaussQuadWeights[l] * refEl.mBasisGQDerivVal[l, i]
                                       * refEl.mBasisGQFunVal[l, j])
            for j, node2 in enumerate(el.mNodesInd):
                for i, node1 in enumerate(el.mNodesInd):
                    dG.append((node1, node2, temp[i, j]))
                    dD.append((node1, node2, temp[i, j]))
    for i, vert in enumerate(mesh.mVertices):
        if isBoundary(vert):
            meshEl = mesh.mFaces[vert.mFaces[0] - 1]
            dgEl = dgMesh.mElements[vert.mFaces[0] - 1]
            if vert.mIndex in bdCond.mDirNodes:
                if vert is meshEl.mVertices[0]:
                    dD.append((dgEl.mNodesInd[e1], dgEl.mNodesInd[e1], 1.0))
                    dC.append((dgEl.mNodesInd[e1], dgEl.mNodesInd[e1], CDir))
                elif vert is meshEl.mVertices[1]:
                    dD.append((dgEl.mNodesInd[e2], dgEl.mNodesInd[e2], -1.0))
                    dC.append((dgEl.mNodesInd[e2], dgEl.mNodesInd[e2], CDir))
                else:
                    raise RuntimeError("vertex / element mismatch")
            elif vert.mIndex in bdCond.mNeuNodes:
                if vert is meshEl.mVertices[0]:
                    dG.append((dgEl.mNodesInd[e1], dgEl.mNodesInd[e1], 1.0))
                elif vert is meshEl.mVertices[1]:
                    dG.append((dgEl.mNodesInd[e2], dgEl.mNodesInd[e2], -1.0))
                else:
                    raise RuntimeError("vertex / element mismatch")
            else:
                raise RuntimeError("Boundary vertex is not included in the boundary condition.")
        else:
            S = dgMesh.mSwitch[i]
            uhatEl = dgMesh.mElements[vert.mFaces[S - 1] - 1]
            qhatEl = dgMesh.mElements[vert.mFaces[S % 2] - 1]
            for k in vert.mFaces:
                meshEl = mesh.mFaces[k - 1]
                dgEl = dgMesh.mElements[k - 1]
                if vert is meshEl.mVertices[0]:
                    dG.append((dgEl.mNodesInd[e1], uhatEl.mNodesInd[e2], 1.0))
                    dD.append((dgEl.mNodesInd[e1], qhatEl.mNodesInd[e1], 1.0))
                elif vert is meshEl.mVertices[1]:
                    dG.append((dgEl.mNodesInd[e2], uhatEl.mNodesInd[e2], -1.0))
                    dD.append((dgEl.mNodesInd[e2], qhatEl.mNodesInd[e1], -1.0))
                else:
                    raise RuntimeError("vertex / element mismatch")
    N = dgMesh.mNumNodes
    return _coo(dG, N), _coo(dD, N), _coo(dC, N)


def _coo(data, N):
    if not data:
        return sp.csc_matrix((N, N))
    r, c, v = zip(*data)
    return sp.csc_matrix((np.array(v, dtype=np.float64), (np.array(r), np.array(c))), shape=(N, N))


def dg_flux_rhs(dgMesh, mesh, func, bdCond, CDir):
    f = np.zeros(dgMesh.mNumNodes)
    r = np.zeros(dgMesh.mNumNodes)
    refEl = dgMesh.mRefEl
    e1 = 0
    e2 = 1 if dgMesh.mP >= 1 else 0
    for el in dgMesh.mElements:
        for i, node in enumerate(el.mNodesInd):
            for l in range(len(refEl.mGaussQuadNodes)):
                f[node] += (el.mJacobian * refEl.mGaussQuadWeights[l] * refEl.mBasisGQFunVal[l, i]
                            * func(el.mRefMap(refEl.mGaussQuadNodes[l])))
    for i, nodeIdx in enumerate(bdCond.mDirNodes):
        vert = mesh.mVertices[nodeIdx - 1]
        dirVal = bdCond.mDirVals[i]
        meshEl = mesh.mFaces[vert.mFaces[0] - 1]
        dgEl = dgMesh.mElements[vert.mFaces[0] - 1]
        if vert is meshEl.mVertices[0]:
            f[dgEl.mNodesInd[e1]] += CDir * dirVal
            r[dgEl.mNodesInd[e1]] += -dirVal
        elif vert is meshEl.mVertices[1]:
            f[dgEl.mNodesInd[e2]] += CDir * dirVal
            r[dgEl.mNodesInd[e2]] += dirVal
        else:
            raise RuntimeError("vertex / element mismatch")
    for nodeIdx in bdCond.mNeuNodes:
        vert = mesh.mVertices[nodeIdx - 1]
        meshEl = mesh.mFaces[vert.mFaces[0] - 1]
        dgEl = dgMesh.mElements[vert.mFaces[0] - 1]
        if vert is meshEl.mVertices[0]:
            f[dgEl.mNodesInd[e1]] += -bdCond.mBdCond[0][1]
        elif vert is meshEl.mVertices[1]:
            f[dgEl.mNodesInd[e2]] += bdCond.mBdCond[1][1]
        else:
            raise RuntimeError("vertex / element mismatch")
    return f, r


def dg_operator_and_rhs(dgMesh, mesh, func, bdCond, CDir):
    """The composition every DG script performs (tests/dg_heirarchy_test.jl:38-42)."""
    G, D, C = dg_flux_operators(dgMesh, mesh, bdCond, CDir)
    A = (C - D @ dgMesh.mMassMatrixLU.solve(G)).tocsc()
    f, r = dg_flux_rhs(dgMesh, mesh, func, bdCond, CDir)
    b = f - D @ dgMesh.mMassMatrixLU.solve(r)
    return A, b, G, D, C
