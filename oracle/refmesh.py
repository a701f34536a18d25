"""Oracle (test infrastructure): 1-D mesh topology, boundary conditions and the mesh generator.

Follows src/meshes.jl:11-69, src/boundary_conditions.jl:1-6, tests/mesh_generator.jl:5-93.

Topology ids (Vertex.mIndex, Face.mIndex, Vertex.mFaces, Face.mNeighbors) keep the reference's
1-based values, including its sentinels (0 = empty slot, -1 / -2 = left / right domain
boundary); python lists are addressed with ``[id - 1]``.  Only DOF index arrays elsewhere in the
oracle are 0-based.
"""


class Vertex:
    """src/meshes.jl:11-17."""

    def __init__(self, mIndex, mX):
        self.mIndex = mIndex
        self.mX = float(mX)
        self.mFaces = [0, 0]


class Face:
    """src/meshes.jl:29-37."""

    def __init__(self, mIndex, nV):
        self.mIndex = mIndex
        self.mVertices = [None] * nV
        self.mNeighbors = [0] * nV


class Mesh:
    """src/meshes.jl:48-51."""

    def __init__(self, vertices, faces):
        self.mVertices = vertices
        self.mFaces = faces


def isBoundary(obj):
    """src/meshes.jl:58-69 (vertex: second face slot < 1; face: last neighbour slot empty)."""
    if isinstance(obj, Face):
        return obj.mNeighbors[-1] == 0
    return obj.mFaces[1] < 1


class BoundaryCondition:
    """src/boundary_conditions.jl:1-6.  mBdCond = [(kind, value) left, (kind, value) right];
    mDirNodes / mNeuNodes hold 1-based *vertex* ids."""

    def __init__(self, mBdCond, mDirNodes, mDirVals, mNeuNodes):
        self.mBdCond = mBdCond
        self.mDirNodes = mDirNodes
        self.mDirVals = mDirVals
        self.mNeuNodes = mNeuNodes


def create_uniform_mesh(n, xin, xout):
    """tests/mesh_generator.jl:5-59.  Vertex i+1 sits at xin + (i/n)*(xout-xin)."""
    faces = [None] * n
    vertices = [None] * (n + 1)
    vertices[0] = Vertex(1, xin)
    for i in range(1, n + 1):
        vertices[i] = Vertex(i + 1, xin + (i / n) * (xout - xin))
        faces[i - 1] = Face(i, 2)
        for j in (1, 2):
            faces[i - 1].mVertices[j - 1] = vertices[i - 1 + j - 1]
    for i in range(n):
        cFace = faces[i]
        for cVertex in cFace.mVertices:
            if cVertex.mFaces[0] == 0:
                cVertex.mFaces[0] = cFace.mIndex
            elif cVertex.mFaces[1] == 0:
                cVertex.mFaces[1] = cFace.mIndex
            else:
                raise RuntimeError("Vertex can only neighbor two faces.")
    for cFace in faces:
        adj = {}
        for fVert in cFace.mVertices:
            for face in fVert.mFaces:
                adj[face] = adj.get(face, 0) + 1
        nIndex = 0
        # The reference iterates a Dict (unordered); neighbour order is never used downstream.
        for f in sorted(adj):
            if f != 0 and f != cFace.mIndex and adj[f] == 1:
                cFace.mNeighbors[nIndex] = f
                nIndex += 1
    return Mesh(vertices, faces)


def set_boundary(mesh, xin, xout, bdCond):
    """tests/mesh_generator.jl:61-93 (``set_boundary!``).  Mutates vertex.mFaces[2] to -1 / -2."""
    dirNodes, dirVals, neuNodes = [], [], []
    for face in mesh.mFaces:
        if isBoundary(face):
            for vert in face.mVertices:
                if isBoundary(vert) and abs(vert.mX - xin) < 1e-15:
                    vert.mFaces[1] = -1
                    if bdCond[0][0] == "dir":
                        dirNodes.append(vert.mIndex)
                        dirVals.append(bdCond[0][1])
                    elif bdCond[0][0] == "neu":
                        neuNodes.append(vert.mIndex)
                elif isBoundary(vert) and abs(vert.mX - xout) < 1e-15:
                    vert.mFaces[1] = -2
                    if bdCond[1][0] == "dir":
                        dirNodes.append(vert.mIndex)
                        dirVals.append(bdCond[1][1])
                    elif bdCond[1][0] == "neu":
                        neuNodes.append(vert.mIndex)
    return BoundaryCondition(list(bdCond), dirNodes, dirVals, neuNodes)
