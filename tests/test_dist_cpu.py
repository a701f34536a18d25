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


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_matches_single_rank(world, tmp_path):
    n = 256
    out = tmp_path / "emu.pkl"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py"), str(out), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    got = pickle.load(open(out, "rb"))
    # single-rank emulation (no communication) ...
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_emulation as emu
    nlev = int(round(math.log2(n)))
    U = uniform.UniformDgHierarchy(n, [3, 1], [2] * nlev, xin=0.0, xout=float(n), CDir=1000.0)
    w = 2.0 * math.pi / 64.0
    b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
    x0 = np.random.default_rng(3).standard_normal(len(b))
    for key, (nPre, nPost) in {"33": (3, 3), "12": (1, 2), "03": (0, 3)}.items():
        ref = emu.vcycle(U, x0, b, 0, 1, nPre=nPre, nPost=nPost)
        assert np.abs(got[key] - ref).max() <= 1e-13 * np.abs(ref).max(), key
    # ... and the emulation itself is the reference V-cycle: compare with the oracle on CSC operators
    from oracle import solvers as osolv
    from oracle.hierarchy import MeshHierarchy
    from oracle.smoother import BlockJacobi
    import scipy.linalg as sla
    S, Sm, I = [], [], []
    for l, lv in enumerate(U.levels):
        lo, di, up = U.level_blocks(l)
        slots = np.arange(lv.n * lv.m).reshape(lv.n, lv.m)
        S.append(blk.blocks_to_csc(lo, di, up, slots, lv.n * lv.m))
        Sm.append(BlockJacobi([sla.lu_factor(d) for d in di], slots.T))
    for l, (P, ratio) in enumerate(U.transfers):
        I.append(uniform._transfer_csc(P, U.levels[l].n, ratio))
    H = MeshHierarchy([None] * len(S), S, None, None, None, Sm, I, None)
    x_or = osolv.multigrid_v_cycle(H, x0, b)
    assert np.abs(got["33"] - x_or).max() <= 1e-11 * np.abs(x_or).max()
