"""Oracle (test infrastructure): the reference's driver-script shapes and the BASELINE configs.

Restates the set-up part of tests/cg_heirarchy_test.jl:11-39, tests/dg_heirarchy_test.jl:11-46,
tests/dg_cg_heirarchy_test.jl:11-50, tests/full_heirarchy_test.jl:13-92 (mesh, boundary
conditions, the p-halving loop ``tempP = div(tempP, 2)``, the hand-built contiguous agglomeration
maps ``(4j-3):(4j)`` / ``(2j-1):(2j)``), and the DG-first + agglomeration shape that BASELINE
configs C2/C3/C5/T need.
"""
import math

import numpy as np

from .aggdg import AgglomeratedDgMesh1, AgglomeratedDgMeshN
from .cg import CgMesh, cg_stiffness_and_rhs
from .dg import DgMesh, dg_operator_and_rhs
from .hierarchy import MeshHierarchyCG, MeshHierarchyDG
from .refmesh import create_uniform_mesh, set_boundary


def agglomeration_maps(n_base, factors):
    """Contiguous 1-based index ranges, one list per agglomerated level
    (tests/full_heirarchy_test.jl:63-75): level 1 groups base elements by factors[0], level i>1
    groups the previous level's elements by factors[i-1]."""
    maps = []
    cur = n_base
    for fac in factors:
        if cur % fac:
            raise ValueError("agglomeration factor does not divide the element count")
        cur //= fac
        maps.append([list(range(fac * j + 1, fac * (j + 1) + 1)) for j in range(cur)])
    return maps


def build_problem(n, cg_orders=(), dg_orders=(), agg_factors=(), pAgg=1, xin=0.0, xout=1.0,
                  CDir=None, func=math.cos, u_exact=math.cos, ux_exact=lambda x: -math.sin(x),
                  bc_kinds=("neu", "dir")):
    """Build (H, x0, b, meta) for a hierarchy CG(cg_orders) -> DG(dg_orders) -> agglomerated."""
    if CDir is None:
        CDir = 1000.0 * n
    mesh = create_uniform_mesh(n, xin, xout)
    vals = []
    for kind, x in zip(bc_kinds, (xin, xout)):
        vals.append((kind, ux_exact(x) if kind == "neu" else u_exact(x)))
    bdCond = set_boundary(mesh, xin, xout, vals)
    nCG, nDG, nAgg = len(cg_orders), len(dg_orders), len(agg_factors)
    meshes = [CgMesh(mesh, p) for p in cg_orders] + [DgMesh(mesh, p) for p in dg_orders]
    if nAgg:
        base = meshes[0] if nCG else meshes[nDG - 1]
        for i, amap in enumerate(agglomeration_maps(n, agg_factors)):
            if i == 0:
                meshes.append(AgglomeratedDgMesh1(pAgg, amap, mesh, base))
            else:
                meshes.append(AgglomeratedDgMeshN(pAgg, amap, meshes[-1], base))
    bdConds = [bdCond] * len(meshes)
    if nCG:
        A, b = cg_stiffness_and_rhs(meshes[0], mesh, func, bdCond)
        H = MeshHierarchyCG(meshes, mesh, bdConds, A, nCG=nCG, nDG=nDG, nAgg=nAgg, CDir=CDir)
    else:
        A, b, G, D, C = dg_operator_and_rhs(meshes[0], mesh, func, bdCond, CDir)
        H = MeshHierarchyDG(meshes, bdConds, A, G, D, C, nDG=nDG, nAgg=nAgg)
    meta = dict(n=n, mesh=mesh, bdCond=bdCond, CDir=CDir, xin=xin, xout=xout)
    return H, np.zeros(len(b)), b, meta


def halving(p, count):
    out = []
    for _ in range(count):
        out.append(p)
        p //= 2
    return out


def cg_heirarchy_test(n=128, maxP=8, nCG=4):
    return build_problem(n, cg_orders=halving(maxP, nCG))


def dg_heirarchy_test(n=128, maxP=8, nDG=4):
    return build_problem(n, dg_orders=halving(maxP, nDG))


def dg_cg_heirarchy_test(n=128, maxP=8, nCG=4, nDG=1):
    orders = halving(maxP, nCG + nDG)
    return build_problem(n, cg_orders=orders[:nCG], dg_orders=orders[nCG:])


def full_heirarchy_test(n=64, maxP=8, pAgg=1, nCG=4):
    nAgg = int(round(math.log2(n))) - 1
    return build_problem(n, cg_orders=halving(maxP, nCG), agg_factors=[4] + [2] * (nAgg - 1),
                         pAgg=pAgg)


def dg_agg_problem(n, p=3, unit_h=False):
    """BASELINE C2/C3/C5/T shape: DG p -> p/2 -> ... -> 1, then pAgg = 1 factor-2 agglomeration
    down to one element.  ``unit_h`` uses the domain [0, n] (h = 1), CDir = 1000 and
    u = cos(2 pi x / 64) so FP64 can reach 1e-10 at large n."""
    orders = []
    q = p
    while q >= 1:
        orders.append(q)
        q //= 2
    nAgg = int(round(math.log2(n)))
    if unit_h:
        w = 2.0 * math.pi / 64.0
        return build_problem(n, dg_orders=orders, agg_factors=[2] * nAgg, xin=0.0, xout=float(n),
                             CDir=1000.0, func=lambda x: w * w * math.cos(w * x),
                             u_exact=lambda x: math.cos(w * x),
                             ux_exact=lambda x: -w * math.sin(w * x))
    return build_problem(n, dg_orders=orders, agg_factors=[2] * nAgg)


def cg_agg_two_level(n=1024):
    """BASELINE C1: CG p=1 on n elements -> one agglomerated level (pAgg = 1, factor 2)."""
    return build_problem(n, cg_orders=[1], agg_factors=[2])
