"""Generates tests/golden/vcycle_golden.npz from the CPU oracle (oracle/, the literal numpy/scipy
restatement of the reference).

The reference (pure Julia + MATLAB.jl plotting) holds no golden vectors and cannot run in this image,
so these fixtures are ORACLE-generated, not reference-generated ("parity unpinned").
They freeze the oracle's behaviour: tests/test_golden.py checks (CPU) that the oracle still reproduces
them and (GPU) that the CUDA path reproduces them through the C ABI.

    python tests/golden/make_golden.py        # rewrites vcycle_golden.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import solvers as osolv          # noqa: E402
from shapes import SHAPES, build_oracle      # noqa: E402

GOLDEN_SHAPES = ["cg_heirarchy", "dg_heirarchy", "dg_cg_heirarchy", "full_heirarchy", "C1_cg1_agg",
                 "C2_dg3_agg", "C3_dg4_agg", "C4_cg3_dg1_agg", "dg_p0_agg0", "factor3_n24"]


def index_digest(H):
    """sha256 over every integer map of the hierarchy's meshes (element-to-DOF, agglomeration maps)."""
    h = hashlib.sha256()
    for m in H.mMeshes:
        for name in ("mNodesInd", "mBaseElementInds", "mSubAggElementInds", "mBlockInds"):
            v = getattr(m, name, None)
            if v is None:
                continue
            if isinstance(v, (list, tuple)):
                for a in v:
                    h.update(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes())
            else:
                h.update(np.ascontiguousarray(np.asarray(v, dtype=np.int64)).tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name in GOLDEN_SHAPES:
        H, x0, b, _ = build_oracle(**SHAPES[name])
        x1 = osolv.multigrid_v_cycle(H, x0, b)
        x, it, res, err = osolv.multigrid(H, x0, b, 100, 1e-10)
        out[f"{name}/b"] = b
        out[f"{name}/x_after_one_vcycle"] = x1
        out[f"{name}/x_final"] = x
        out[f"{name}/iters"] = np.array([it])
        out[f"{name}/res"] = res
        out[f"{name}/err"] = err
        out[f"{name}/index_sha256"] = np.frombuffer(bytes.fromhex(index_digest(H)), dtype=np.uint8)
        print(name, it, res[-1])
    np.savez_compressed(os.path.join(HERE, "vcycle_golden.npz"), **out)


if __name__ == "__main__":
    main()
