"""Scratch diagnostic: long residual histories at 2^26 (C4 generic tier vs fused; T for 40 cycles)."""
import math, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agglomerationmultigrid1d_b200 import uniform
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 2 ** log2n
w = 2.0 * math.pi / 64.0
np.set_printoptions(linewidth=200, precision=3)
U = uniform.UniformDgHierarchy(n, [3, 1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
dev = U.upload()
b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
x, it, res, _ = dev.solve(np.zeros(len(b)), b, 40, 1e-30)
print("T", it, res / np.linalg.norm(b), flush=True)
dev.close()
del U, dev, b, x
U = uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * log2n, xin=0.0, xout=float(n), CDir=1000.0)
dev = U.upload()
b = U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])
for fused, tail in ((1, 1024), (0, 0)):
    dev.set_option("fused", fused)
    dev.set_option("coarse_cta_elems", tail)
    x, it, res, _ = dev.solve(np.zeros(len(b)), b, 24, 1e-30)
    print("C4 fused", fused, it, res / np.linalg.norm(b), flush=True)
