"""Worker for the multi-GPU parity test (run under torch.distributed.run, one rank per GPU).

Each rank builds the same uniform hierarchy, uploads its slab through amg1d_create_dist and solves;
rank 0 also solves the whole problem on a single-GPU handle.  The sharded run must give the same
V-cycle count, the same residual history (up to the order of the norm's final sum) and - because the
per-element arithmetic is identical - bit-identical solution slabs."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from agglomerationmultigrid1d_b200 import _capi as capi, uniform   # noqa: E402


def main():
    out = sys.argv[1]
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    kinds = (sys.argv[3] if len(sys.argv) > 3 else "dg").split(",")
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.cuda.set_device(rank)
    reports = {}
    for kind in kinds:                      # one rendezvous, one NCCL communicator per kind
        reports[kind] = run_kind(kind, log2n)
        dist.barrier()
    if rank == 0:
        json.dump(reports, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def run_graded(world, rank, ids):
    """DG 3 -> 1 -> ten agglomerated levels on a GRADED (non-uniform) mesh of 2^14 elements, host-built by the
    general path (MeshHierarchy: global sparse algebra, explicit per-element operator and transfer blocks - no
    pattern upload anywhere): every rank passes the same global arrays and keeps its slab.  Bit-identical to the
    single-GPU handle."""
    import agglomerationmultigrid1d_b200 as aggmg
    n = 2 ** 14
    rng = np.random.default_rng(0)
    x = np.concatenate([[0.0], np.cumsum(0.5 + rng.random(n))])
    xout = float(x[-1])
    w = 2.0 * math.pi / 64.0
    mesh = aggmg.Mesh(x)
    bd = aggmg.set_boundary(mesh, 0.0, xout, [("neu", 0.0), ("dir", math.cos(w * xout))])
    meshes = [aggmg.DgMesh(mesh, 3), aggmg.DgMesh(mesh, 1)]
    cur = n
    for i in range(10):
        agg = [[2 * j, 2 * j + 1] for j in range(cur // 2)]
        cur //= 2
        meshes.append(aggmg.AgglomeratedDgMesh1(1, agg, mesh, meshes[1]) if i == 0
                      else aggmg.AgglomeratedDgMeshN(1, agg, meshes[-1], meshes[1]))
    G, D, C = aggmg.dg_flux_operators(meshes[0], mesh, bd, 1000.0)
    A = (C - D @ meshes[0].mMassMatrixLU.solve(G)).tocsc()
    f, r = aggmg.dg_flux_rhs(meshes[0], mesh, lambda t: w * w * np.cos(w * t), bd, 1000.0)
    b = f - D @ meshes[0].mMassMatrixLU.solve(r)
    H = aggmg.MeshHierarchy(meshes, [bd] * len(meshes), A, G, D, C, nDG=2, nAgg=10, upload=False)
    dev = H.upload(device=rank, dist=(rank, world, ids[0]), options={"shard_min": 256, "leg_pipeline_min": 0})   # pipelined legs on the slabs
    nloc = n // world
    lo, hi = rank * nloc * 4, (rank + 1) * nloc * 4
    report = {"rank": rank, "gather_level": dev.info("gather_level"), "local_dofs": dev.info("local_dofs"),
              "ghost_depth": dev.info("ghost_depth"), "p2p_halo": dev.info("p2p_halo"), "p2p_requested": 1}
    x_loc, it, res, _ = dev.solve(np.zeros(hi - lo), b[lo:hi], 100, 1e-10)
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(len(b))
    v = dev.vcycle(x0[lo:hi], b[lo:hi], nPre=2, nPost=1, alpha=0.7)
    gathered = [None] * world
    dist.gather_object((x_loc, it, res, v), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        d1 = H.upload(device=0)
        x1, it1, res1, _ = d1.solve(np.zeros(len(b)), b, 100, 1e-10)
        v1 = d1.vcycle(x0, b, nPre=2, nPost=1, alpha=0.7)
        ok, msgs = True, []
        if gathered[0][1] != it1 or not np.allclose(gathered[0][2], res1, rtol=1e-12, atol=0):
            ok = False; msgs.append(f"iters / history {gathered[0][1]} vs {it1}")
        if not np.array_equal(np.concatenate([g[0] for g in gathered]), x1):
            ok = False; msgs.append("solve x differs")
        if not np.array_equal(np.concatenate([g[3] for g in gathered]), v1):
            ok = False; msgs.append("vcycle x differs")
        if not res1[-1] < 1e-10 * np.linalg.norm(b):
            ok = False; msgs.append("not converged")
        report.update(ok=ok, msgs=msgs, iters=int(it1), world=world, res_last=float(res1[-1]))
        d1.close()
    dev.close()
    return report


def run_kind(kind, log2n):
    if kind == "dg_graded":
        lib = capi.load()
        ids = [None]
        if dist.get_rank() == 0:
            buf = C.create_string_buffer(128)
            capi.check(None, lib.amg1d_nccl_unique_id(C.cast(buf, C.c_void_p)))
            ids[0] = buf.raw
        dist.broadcast_object_list(ids, src=0)
        return run_graded(dist.get_world_size(), dist.get_rank(), ids)
    # *_pat / *_pat2: the sharded handle reads its operators from the pattern tables (option pattern_resident
    # = 1) / takes the interior block set as constant-bank kernel parameters (= 2)
    pattern_resident = 2 if kind.endswith("_pat2") else 1 if kind.endswith("_pat") else 0
    kind = kind[:-5] if pattern_resident == 2 else kind[:-4] if pattern_resident else kind
    # *_nccl: the slab edges travel through NCCL send / recv groups instead of peer memory (option p2p_halo = 0)
    p2p = 0 if kind.endswith("_nccl") else 1
    kind = kind[:-5] if not p2p else kind
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = capi.load()
    ids = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        capi.check(None, lib.amg1d_nccl_unique_id(C.cast(buf, C.c_void_p)))
        ids[0] = buf.raw
    dist.broadcast_object_list(ids, src=0)
    n = 2 ** log2n
    w = 2.0 * math.pi / 64.0
    func = lambda x: w * w * np.cos(w * x)                       # noqa: E731
    vals = [0.0, math.cos(w * n)]

    def build():
        if kind == "cg":        # BASELINE C4 shape: CG 3 -> 1 -> DG 1 -> agglomerated levels
            return uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
        if kind == "cg8":       # the order of the reference's CG scripts: 8 x 8 groups, row-per-thread legs
            return uniform.UniformCgHierarchy(n, [8, 4, 2, 1], [1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
        if kind == "dg8":       # tests/dg_heirarchy_test.jl order: 9 x 9 blocks, row-per-thread legs
            return uniform.UniformDgHierarchy(n, [8, 4, 2, 1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
        return uniform.UniformDgHierarchy(n, [3, 1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)

    U = build()
    dev = U.upload(device=rank, dist=(rank, world, ids[0]), options={"shard_min": 512, "p2p_halo": p2p,
                                                                     "leg_pipeline_min": 0})   # persistent pipelined legs
    # on the sharded 4 x 4 levels even at this size (the single-GPU reference below keeps the default: one window per CTA)
    if pattern_resident:                          # the single-GPU reference below keeps streaming its operators
        dev.set_option("pattern_resident", pattern_resident)
    nloc = n // world
    m0 = U.levels[0].m
    is_cg = kind.startswith("cg")
    if is_cg:                   # slabs of groups; the last rank also holds the closing vertex group
        lo, hi = rank * nloc, (rank + 1) * nloc + (1 if rank == world - 1 else 0)
        b_loc = U.rhs(func, vals, group_range=(lo, hi))
        n_blocks = n + 1
    else:
        lo, hi = rank * nloc, (rank + 1) * nloc
        b_loc = U.rhs(func, vals, elem_range=(lo, hi))
        n_blocks = n
    report = {"rank": rank, "gather_level": dev.info("gather_level"), "local_dofs": dev.info("local_dofs"),
              "ghost_depth": dev.info("ghost_depth"), "p2p_halo": dev.info("p2p_halo"), "p2p_requested": p2p}
    results = {}
    # right-hand side assembled on the device, every rank its own slab (+ ghosts): bit-identical to the single-GPU
    # assembly, and the solve on the device-resident problem takes the same number of cycles
    terms = [("cos", w * w, 0, w, 0.0)]
    U.device_rhs(terms, vals, dev)
    b_dev = dev.dev_get_rhs()
    it_dev, res_dev = dev.dev_solve(100, 1e-10)
    results["device_rhs"] = (b_dev, it_dev, res_dev)
    x, it, res, _ = dev.solve(np.zeros(len(b_loc)), b_loc, 100, 1e-10)
    results["solve"] = (x, it, res)
    rng = np.random.default_rng(5)
    x0_glob = rng.standard_normal(n_blocks * m0)
    if is_cg:
        x0_glob.reshape(n_blocks, m0)[n, 1:] = 0.0               # padding slots of the closing group
    x0_loc = x0_glob[lo * m0:hi * m0]
    for key, (nPre, nPost, alpha) in {"v312": (3, 1, 0.5), "v023": (0, 2, 2.0 / 3.0), "v330": (3, 3, 0.8)}.items():
        results[key] = (dev.vcycle(x0_loc, b_loc, nPre=nPre, nPost=nPost, alpha=alpha), 0, np.zeros(0))
    gathered = [None] * world
    dist.gather_object({k: v for k, v in results.items()}, gathered if rank == 0 else None, dst=0)
    ok, msgs = True, []
    if rank == 0:
        U1 = build()
        d1 = U1.upload(device=0)
        b = U1.rhs(func, vals)
        x1, it1, res1, _ = d1.solve(np.zeros(len(b)), b, 100, 1e-10)
        xs = np.concatenate([g["solve"][0] for g in gathered])
        it_d, res_d = gathered[0]["solve"][1], gathered[0]["solve"][2]
        if it_d != it1:
            ok = False; msgs.append(f"iters {it_d} vs {it1}")
        elif not np.allclose(res_d, res1, rtol=1e-12, atol=0):
            ok = False; msgs.append(f"res {res_d} vs {res1}")
        if not np.array_equal(xs, x1):
            ok = False; msgs.append(f"solve x max diff {np.abs(xs - x1).max():.3e}")
        U1.device_rhs(terms, vals, d1)
        b1 = d1.dev_get_rhs()
        it1d, res1d = d1.dev_solve(100, 1e-10)
        bs = np.concatenate([g["device_rhs"][0] for g in gathered])
        if not np.array_equal(bs, b1):
            ok = False; msgs.append(f"device rhs max diff {np.abs(bs - b1).max():.3e}")
        if np.abs(b1 - b).max() > 1e-13 * np.abs(b).max():
            ok = False; msgs.append(f"device rhs vs host rhs {np.abs(b1 - b).max():.3e}")
        if gathered[0]["device_rhs"][1] != it1d or not np.allclose(gathered[0]["device_rhs"][2], res1d, rtol=1e-12, atol=0):
            ok = False; msgs.append(f"device-rhs solve: iters {gathered[0]['device_rhs'][1]} vs {it1d}")
        for key, (nPre, nPost, alpha) in {"v312": (3, 1, 0.5), "v023": (0, 2, 2.0 / 3.0), "v330": (3, 3, 0.8)}.items():
            ref = d1.vcycle(x0_glob, b, nPre=nPre, nPost=nPost, alpha=alpha)
            got = np.concatenate([g[key][0] for g in gathered])
            if not np.array_equal(got, ref):
                ok = False; msgs.append(f"{key} x max diff {np.abs(got - ref).max():.3e}")
        report.update(ok=ok, msgs=msgs, iters=int(it1), world=world, res_last=float(res1[-1]))
        d1.close()
    dev.close()
    return report


if __name__ == "__main__":
    main()
