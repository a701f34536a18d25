"""gloo worker for tests/test_dist_cpu.py: world_size ranks emulate the sharded V-cycle on CPU."""
import math
import os
import pickle
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from agglomerationmultigrid1d_b200 import uniform        # noqa: E402
import dist_emulation as emu                              # noqa: E402


def main():
    out, n = sys.argv[1], int(sys.argv[2])
    kind = sys.argv[3] if len(sys.argv) > 3 else "dg"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nlev = int(round(math.log2(n)))
    w = 2.0 * math.pi / 64.0
    func, vals = (lambda x: w * w * np.cos(w * x)), [0.0, math.cos(w * n)]
    nloc = n // world
    if kind == "cg":      # BASELINE C4 shape: slabs of vertex groups, closing group on the last rank
        U = uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * nlev, xin=0.0, xout=float(n), CDir=1000.0)
        b = U.rhs(func, vals)
        lo, hi = rank * nloc, (rank + 1) * nloc + (1 if rank == world - 1 else 0)
        b_loc = U.rhs(func, vals, group_range=(lo, hi))
        gd = 4 + 1
    else:
        U = uniform.UniformDgHierarchy(n, [3, 1], [2] * nlev, xin=0.0, xout=float(n), CDir=1000.0)
        b = U.rhs(func, vals)
        lo, hi = rank * nloc, (rank + 1) * nloc
        b_loc = U.rhs(func, vals, elem_range=(lo, hi))
        gd = 4
    m0 = U.levels[0].m
    # each rank assembles only its own slab of the right-hand side
    assert np.array_equal(b_loc, b[lo * m0:hi * m0])
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(len(b))
    if kind == "cg":
        x0.reshape(-1, m0)[n, 1:] = 0.0                   # padding slots of the closing group
    res = {}
    for key, (nPre, nPost) in {"33": (3, 3), "12": (1, 2), "03": (0, 3)}.items():
        res[key] = emu.vcycle(U, x0, b, rank, world, nPre=nPre, nPost=nPost, gd=gd)
    parts = [None] * world
    dist.gather_object(res, parts if rank == 0 else None, dst=0)
    if rank == 0:
        pickle.dump({k: np.concatenate([p[k] for p in parts]) for k in res}, open(out, "wb"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
