"""Scratch diagnostic: convergence of the C4 hierarchy (CG 3 -> 1 -> DG 1 -> agglomerated) vs n."""
import math
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agglomerationmultigrid1d_b200 import uniform

for log2n in [int(a) for a in sys.argv[1:]] or [22, 23, 24]:
    n = 2 ** log2n
    w = 2.0 * math.pi / 64.0
    U = uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
    dev = U.upload()
    b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
    for fused in (1, 0):
        dev.set_option("fused", fused)
        x, it, res, _ = dev.solve(np.zeros(len(b)), b, 30, 1e-10)
        print(log2n, "fused", fused, "iters", it, "res/|b|", (res / np.linalg.norm(b))[:12], flush=True)
        if log2n >= 25:
            break
    dev.close()
