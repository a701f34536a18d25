#!/usr/bin/env python
"""Residual history of `multigrid` (src/solvers.jl:116-139) on one of bench.py's workloads at a given size,
from the CPU oracle (oracle/vcycle_ref.c in block-pattern storage - runs anywhere, no GPU) or from the GPU
library, for a FIXED number of cycles (no stopping test), written as JSON.  Used to put the two histories
side by side at the sizes where FP64 stops converging.

  python tools/history.py --impl oracle --workload C4 --log2n 26 --cycles 25 --out profiles/r02_hist_oracle_C4_26.json
  python tools/history.py --impl gpu    --workload C4 --log2n 26 --cycles 25 --out gpurun_out/hist_gpu_C4_26.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "gpu"], required=True)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--cycles", type=int, default=25)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    n = 2 ** a.log2n
    pr = bench.problem(n)
    t0 = time.perf_counter()
    U = bench.build_hierarchy(a.workload, n)
    b = U.rhs(pr["func"], pr["bc_values"])
    nb = float(np.linalg.norm(b))
    t_setup = time.perf_counter() - t0
    rec = {"impl": a.impl, "workload": a.workload, "log2n": a.log2n, "n_elements": n, "fine_dofs": len(b),
           "levels": len(U.levels), "rhs_norm": nb, "cycles": a.cycles, "setup_s": t_setup}
    t0 = time.perf_counter()
    if a.impl == "oracle":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import cref
        c = cref.CRefPattern(*cref.pattern_arrays(U))
        rec["threads"] = c.set_threads(a.threads)
        x = np.zeros(len(b))
        res = []
        for i in range(a.cycles):
            x = c.vcycle(x, b)
            res.append(c.residual_norm(x, b))
            print(f"cycle {i + 1}: ||Ax-b||/||b|| = {res[-1] / nb:.6e}", flush=True)
        c.close()
    else:
        dev = U.upload()
        x, it, res, _ = dev.solve(np.zeros(len(b)), b, a.cycles, 0.0)        # tol = 0: never stops early
        res = list(map(float, res))
        dev.close()
    rec["seconds"] = time.perf_counter() - t0
    rec["res"] = [float(r) for r in res]
    rec["rel"] = [float(r) / nb for r in res]
    rec["x_absmax"] = float(np.abs(x).max())
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(rec, open(a.out, "w"), indent=1)
    print(json.dumps({k: rec[k] for k in ("impl", "workload", "log2n", "seconds")}), flush=True)


if __name__ == "__main__":
    main()
