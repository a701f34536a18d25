"""Problem shapes shared by the tests: each is built twice, independently - by the oracle's literal
restatement (oracle/drivers.py) and by the product's host package - from the same parameters."""
import math

import numpy as np

import agglomerationmultigrid1d_b200 as aggmg
from oracle import drivers


def halving(p, count):
    return drivers.halving(p, count)


def dg_orders_to_one(p):
    out = []
    while p >= 1:
        out.append(p)
        p //= 2
    return out


# name -> kwargs of build (same keys for both builders)
SHAPES = {
    # the four reference driver scripts (tests/*_heirarchy_test.jl) at a size the oracle does in seconds
    "cg_heirarchy": dict(n=32, cg_orders=halving(8, 4)),
    "dg_heirarchy": dict(n=32, dg_orders=halving(8, 4)),
    "dg_cg_heirarchy": dict(n=32, cg_orders=halving(8, 4), dg_orders=[0]),
    "full_heirarchy": dict(n=32, cg_orders=halving(8, 4), agg_factors=[4, 2, 2, 2]),
    # BASELINE configs, scaled down
    "C1_cg1_agg": dict(n=64, cg_orders=[1], agg_factors=[2]),
    "C1_cg1_agg_n1024": dict(n=1024, cg_orders=[1], agg_factors=[2]),        # BASELINE configs[0] at its own size
    "C2_dg3_agg": dict(n=64, dg_orders=[3, 1], agg_factors=[2] * 6),
    "C2_dg3_agg_unit_h": dict(n=64, dg_orders=[3, 1], agg_factors=[2] * 6, unit_h=True),
    "C3_dg4_agg": dict(n=32, dg_orders=[4, 2, 1], agg_factors=[2] * 5),
    "C4_cg3_dg1_agg": dict(n=32, cg_orders=[3, 1], dg_orders=[1], agg_factors=[2] * 5),
    "dg_p0_agg0": dict(n=16, dg_orders=[1], agg_factors=[2, 2], pAgg=0),
    # coarsest levels with more than 64 elements: solved by block cyclic reduction on the GPU
    "bcr_dg_cg_n128": dict(n=128, cg_orders=[4, 2, 1], dg_orders=[0]),       # literal dg_cg_heirarchy: size-n coarse solve
    "bcr_dg3_agg_n512": dict(n=512, dg_orders=[3, 1], agg_factors=[2]),     # coarsest: 256 agglomerated elements
    "factor3_n24": dict(n=24, dg_orders=[2, 1], agg_factors=[3, 2, 2, 2]),
}


def _problem(unit_h, n):
    if unit_h:
        w = 2.0 * math.pi / 64.0
        return dict(xin=0.0, xout=float(n), CDir=1000.0,
                    func=lambda x: w * w * np.cos(w * x), u=lambda x: np.cos(w * x),
                    ux=lambda x: -w * np.sin(w * x))
    return dict(xin=0.0, xout=1.0, CDir=1000.0 * n, func=np.cos, u=np.cos, ux=lambda x: -np.sin(x))


def build_oracle(n, cg_orders=(), dg_orders=(), agg_factors=(), pAgg=1, unit_h=False):
    pr = _problem(unit_h, n)
    return drivers.build_problem(
        n, cg_orders=cg_orders, dg_orders=dg_orders, agg_factors=agg_factors, pAgg=pAgg,
        xin=pr["xin"], xout=pr["xout"], CDir=pr["CDir"], func=lambda x: float(pr["func"](x)),
        u_exact=lambda x: float(pr["u"](x)), ux_exact=lambda x: float(pr["ux"](x)))


def build_package(n, cg_orders=(), dg_orders=(), agg_factors=(), pAgg=1, unit_h=False, upload=True,
                  bc_kinds=("neu", "dir"), device_setup=False):
    """The product's host path, written the way the reference scripts are."""
    pr = _problem(unit_h, n)
    xin, xout, CDir = pr["xin"], pr["xout"], pr["CDir"]
    mesh = aggmg.create_uniform_mesh(n, xin, xout)
    vals = [(k, float(pr["ux"](x)) if k == "neu" else float(pr["u"](x)))
            for k, x in zip(bc_kinds, (xin, xout))]
    bdCond = aggmg.set_boundary(mesh, xin, xout, vals)
    meshes = [aggmg.CgMesh(mesh, p) for p in cg_orders] + [aggmg.DgMesh(mesh, p) for p in dg_orders]
    nCG, nDG, nAgg = len(cg_orders), len(dg_orders), len(agg_factors)
    if nAgg:
        base = meshes[0] if nCG else meshes[nDG - 1]
        cur = n
        for i, fac in enumerate(agg_factors):
            agg = [list(range(fac * j, fac * (j + 1))) for j in range(cur // fac)]
            cur //= fac
            if i == 0:
                meshes.append(aggmg.AgglomeratedDgMesh1(pAgg, agg, mesh, base))
            else:
                meshes.append(aggmg.AgglomeratedDgMeshN(pAgg, agg, meshes[-1], base))
    bdConds = [bdCond] * len(meshes)
    if nCG:
        A, b = aggmg.cg_stiffness_and_rhs(meshes[0], mesh, pr["func"], bdCond)
        H = aggmg.MeshHierarchy(meshes, mesh, bdConds, A, nCG=nCG, nDG=nDG, nAgg=nAgg, CDir=CDir,
                                upload=upload, device_setup=device_setup)
    else:
        G, D, C = aggmg.dg_flux_operators(meshes[0], mesh, bdCond, CDir)
        A = (C - D @ meshes[0].mMassMatrixLU.solve(G)).tocsc()
        f, r = aggmg.dg_flux_rhs(meshes[0], mesh, pr["func"], bdCond, CDir)
        b = f - D @ meshes[0].mMassMatrixLU.solve(r)
        H = aggmg.MeshHierarchy(meshes, bdConds, A, G, D, C, nDG=nDG, nAgg=nAgg, upload=upload,
                                device_setup=device_setup)
    return H, np.zeros(len(b)), b
