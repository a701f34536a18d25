"""Host set-up: 1-D mesh topology, boundary conditions, mesh generator (array form).

Mirrors src/meshes.jl:11-69, src/boundary_conditions.jl:1-6 and tests/mesh_generator.jl:5-93.  The
reference builds Vertex / Face object graphs; a 1-D mesh is fully described by its sorted vertex
coordinates, so the same information is held as arrays: face k (0-based) joins vertices k, k+1.
"""
import numpy as np


class Mesh:
    """``mVertexX[v]``: coordinate of vertex v; ``mBdSide[v]``: 0 interior, -1 / -2 the left / right
    domain boundary once ``set_boundary`` ran (the reference stores these in ``mFaces[2]``)."""

    def __init__(self, vertex_x):
        self.mVertexX = np.ascontiguousarray(vertex_x, dtype=np.float64)
        if self.mVertexX.ndim != 1 or len(self.mVertexX) < 2:
            raise ValueError("a mesh needs at least two vertices")
        if not np.all(np.diff(self.mVertexX) > 0):
            raise ValueError("vertex coordinates must be strictly increasing")
        self.mBdSide = np.zeros(len(self.mVertexX), dtype=np.int64)

    @property
    def nFaces(self):
        return len(self.mVertexX) - 1

    @property
    def nVertices(self):
        return len(self.mVertexX)

    def isBoundary(self, v):
        """src/meshes.jl:58-60: a vertex that belongs to a single face."""
        return v == 0 or v == self.nVertices - 1


class BoundaryCondition:
    """src/boundary_conditions.jl:1-6.  mBdCond = [(kind, value) at xin, (kind, value) at xout] with
    kind 'dir' or 'neu'; node lists hold 0-based vertex ids in the reference's traversal order."""

    def __init__(self, mBdCond, mDirNodes, mDirVals, mNeuNodes):
        self.mBdCond = list(mBdCond)
        self.mDirNodes = list(mDirNodes)
        self.mDirVals = list(mDirVals)
        self.mNeuNodes = list(mNeuNodes)

    def kind(self, side):
        """side 0 = left (xin), 1 = right (xout)."""
        return self.mBdCond[side][0]

    def value(self, side):
        return self.mBdCond[side][1]


def create_uniform_mesh(n, xin, xout):
    """tests/mesh_generator.jl:5-59; vertex i at xin + (i/n)*(xout-xin) (note: not i*h)."""
    n = int(n)
    if n < 1:
        raise ValueError("n must be >= 1")
    i = np.arange(n + 1, dtype=np.float64)
    x = xin + (i / n) * (xout - xin)
    x[0] = xin
    return Mesh(x)


def set_boundary(mesh, xin, xout, bdCond):
    """``set_boundary!`` (tests/mesh_generator.jl:61-93): tags the two end vertices (tolerance
    1e-15 on |x - xin|, |x - xout|) and returns the BoundaryCondition."""
    for kind, _ in bdCond:
        if kind not in ("dir", "neu"):
            raise ValueError("boundary kind must be 'dir' or 'neu'")
    dirNodes, dirVals, neuNodes = [], [], []
    ends = [0, mesh.nVertices - 1] if mesh.nFaces > 1 else [0, 1]
    for v in ends:
        x = mesh.mVertexX[v]
        if abs(x - xin) < 1e-15:
            side = 0
        elif abs(x - xout) < 1e-15:
            side = 1
        else:
            continue
        mesh.mBdSide[v] = -1 - side
        if bdCond[side][0] == "dir":
            dirNodes.append(v)
            dirVals.append(bdCond[side][1])
        else:
            neuNodes.append(v)
    return BoundaryCondition(bdCond, dirNodes, dirVals, neuNodes)
