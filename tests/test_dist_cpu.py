"""N > 1 host-side logic on CPU (gloo, world_size 2 and 4): slab plan, per-rank right-hand-side slabs,
and the halo / gather / scatter schedule of the sharded V-cycle (emulated with numpy blocks) against
the single-rank run and against the oracle."""
import math
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from agglomerationmultigrid1d_b200 import slabs, uniform
from agglomerationmultigrid1d_b200 import blocks as blk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_slabs():
    sizes = [2 ** 20, 2 ** 20] + [2 ** k for k in range(19, -1, -1)]
    ratios = [1] + [2] * 20
    plan, g = slabs.plan_slabs(sizes, ratios, rank=3, nranks=8, shard_min=8192)
    assert g == sizes.index(2 ** 15)                      # first level with < 8 * 8192 elements
    assert all(p.sharded for p in plan[:g]) and not any(p.sharded for p in plan[g:])
    assert plan[0].n == 2 ** 17 and plan[0].start == 3 * 2 ** 17 and plan[0].gl == 4 and plan[0].gr == 4
    assert not plan[g].present
    p0, _ = slabs.plan_slabs(sizes, ratios, rank=0, nranks=8)
    assert p0[0].gl == 0 and p0[0].gr == 4 and p0[g].present
    plan1, g1 = slabs.plan_slabs(sizes, ratios, rank=0, nranks=1)
    assert g1 == -1 and not any(p.sharded for p in plan1)
    with pytest.raises(ValueError):
        slabs.plan_slabs([100, 50], [2], rank=0, nranks=8)             # finest level not shardable
    # slab edges stay aligned to the agglomerates of every sharded level
    for l in range(g - 1):
        assert plan[l].start % ratios[l] == 0 and plan[l].start // ratios[l] == plan[l + 1].start


def _oracle_from_uniform(U):
    """Oracle MeshHierarchy (CSC operators, LU blocks / point Jacobi, sparse transfers) of a uniform
    pattern hierarchy in its device ordering (group order for CG levels)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_emulation as emu
    from oracle.hierarchy import MeshHierarchy
    from oracle.smoother import BlockJacobi, JacobiSmoother
    import scipy.linalg as sla
    S, Sm, I = [], [], []
    for l, lv in enumerate(U.levels):
        lo, di, up = U.level_blocks(l)
        slots = np.arange(lv.n * lv.m).reshape(lv.n, lv.m)
        S.append(blk.blocks_to_csc(lo, di, up, slots, lv.n * lv.m))
        if getattr(lv, "is_cg", False):
            Sm.append(JacobiSmoother(S[-1].diagonal()))
        else:
            Sm.append(BlockJacobi([sla.lu_factor(d) for d in di], slots.T))
    for l in range(len(U.levels) - 1):
        parent, P0, P1 = emu.transfer_of(U, l)
        nf, mf, mc = P0.shape
        nc = U.levels[l + 1].n
        rows, cols, vals = [], [], []
        for P, off in ((P0, 0), (P1, 1)):
            if P is None:
                continue
            par = parent + off
            e = np.flatnonzero((par >= 0) & (par < nc))
            rows.append(np.repeat((e[:, None] * mf + np.arange(mf))[:, :, None], mc, axis=2).ravel())
            cols.append(np.repeat((par[e][:, None] * mc + np.arange(mc))[:, None, :], mf, axis=1).ravel())
            vals.append(P[e].ravel())
        I.append(sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                               shape=(nf * mf, nc * mc)))
    return MeshHierarchy([None] * len(S), S, None, None, None, Sm, I, None)


@pytest.mark.parametrize("kind", ["dg", "cg"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_matches_single_rank(world, kind, tmp_path):
    n = 256
    out = tmp_path / "emu.pkl"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + (10 if kind == "cg" else 0)),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py"), str(out), str(n), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    got = pickle.load(open(out, "rb"))
    # single-rank emulation (no communication) ...
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_emulation as emu
    nlev = int(round(math.log2(n)))
    w = 2.0 * math.pi / 64.0
    if kind == "cg":
        U = uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * nlev, xin=0.0, xout=float(n), CDir=1000.0)
    else:
        U = uniform.UniformDgHierarchy(n, [3, 1], [2] * nlev, xin=0.0, xout=float(n), CDir=1000.0)
    b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
    x0 = np.random.default_rng(3).standard_normal(len(b))
    if kind == "cg":
        x0.reshape(-1, U.levels[0].m)[n, 1:] = 0.0
    for key, (nPre, nPost) in {"33": (3, 3), "12": (1, 2), "03": (0, 3)}.items():
        ref = emu.vcycle(U, x0, b, 0, 1, nPre=nPre, nPost=nPost)
        assert np.abs(got[key] - ref).max() <= 1e-13 * np.abs(ref).max(), key
    # ... and the emulation itself is the reference V-cycle: compare with the oracle on CSC operators
    from oracle import solvers as osolv
    H = _oracle_from_uniform(U)
    x_or = osolv.multigrid_v_cycle(H, x0, b)
    assert np.abs(got["33"] - x_or).max() <= 1e-11 * np.abs(x_or).max()


def test_plan_slabs_cg_groups():
    """CG levels have n + 1 vertex groups: the closing group goes to the last rank, and the slab
    starts of the CG levels coincide with those of the DG level underneath (group k <-> element k)."""
    sizes = [257, 257, 256, 128, 64, 32, 16, 8]
    ratios = [1, 1, 2, 2, 2, 2, 2]
    plan, g = slabs.plan_slabs(sizes, ratios, rank=1, nranks=2, shard_min=16, ghost_depth=5)
    assert g == sizes.index(16)
    assert plan[0].sharded and plan[0].n == 129 and plan[0].start == 128
    assert plan[1].n == 129 and plan[2].n == 128 and plan[2].start == 128
    p0, _ = slabs.plan_slabs(sizes, ratios, rank=0, nranks=2, shard_min=16, ghost_depth=5)
    assert p0[0].n == 128 and p0[0].gr == 5 and p0[0].gl == 0 and p0[g].present and not plan[g].present
    assert slabs.slab_size(257, 2, 0) + slabs.slab_size(257, 2, 1) == 257
