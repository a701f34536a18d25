"""Small end-to-end exercise of every kernel family, sized for compute-sanitizer.

  compute-sanitizer --tool memcheck  --error-exitcode 3 python tools/sanitize.py
  compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py

(compute-sanitizer is closed on the GPU pool this repository was developed on - profiles/r01d_ncu_summary.md; the
script also runs plainly, as a self-check of every fast path against the generic tier.)

Covers: the fused legs f_down / f_up (1x1 .. 5x5 blocks, block and point Jacobi, one- and two-parent
transfers), the row-per-thread legs r_down / r_up (9x9 DG blocks, CG groups of 8; 1 and 3 rows per thread),
the single-CTA tail f_tail, pattern-resident operators, the generic tier (fused = 0), block cyclic reduction,
PCG, ldiv, the device-side set-up chain.  Every run is compared with the generic tier (bit-identical x), so
a sanitizer-clean exit also means the numbers were right under the tool."""
import math
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rhs(U, n, unit_h=True):
    w = 2.0 * math.pi / 64.0
    return U.rhs(lambda x: w * w * np.cos(w * x), [0.0, math.cos(w * n)])


def exercise(U, n, label, variants):
    dev = U.upload()
    try:
        b = rhs(U, n)
        x0 = np.random.default_rng(3).standard_normal(len(b))
        dev.set_option("fused", 0)
        ref = dev.vcycle(x0, b)
        dev.set_option("fused", 1)
        for opts in variants:
            for k, v in opts.items():
                dev.set_option(k, v)
            x = dev.vcycle(x0, b)
            assert np.array_equal(x, ref), (label, opts)
            x2 = dev.vcycle(x0, b, nPre=1, nPost=2, alpha=0.7)
            dev.set_option("fused", 0)
            assert np.array_equal(x2, dev.vcycle(x0, b, nPre=1, nPost=2, alpha=0.7)), (label, opts)
            dev.set_option("fused", 1)
        xs, it, res, _ = dev.solve(np.zeros(len(b)), b, 30, 1e-10)
        xp, itp, resp = dev.pcg(np.zeros(len(b)), b, 30, 1e-10)[:3]
        y = dev.ldiv(b)
        assert np.array_equal(y, dev.vcycle(np.zeros(len(b)), b))
        u = dev.direct_solve(0, b)
        print(f"[sanitize] {label}: {it} V-cycles, {itp} PCG iterations, launches per cycle "
              f"{dev.info('launches_per_cycle')}, |x - A\\b|_max / |x|_max = "
              f"{np.abs(u - xs).max() / max(1e-300, np.abs(u).max()):.2e}", flush=True)
    finally:
        dev.close()


def main():
    from agglomerationmultigrid1d_b200 import uniform
    pat = [{"pattern_resident": 0}, {"pattern_resident": 1}, {"pattern_resident": 2}]
    n = 1536                                                     # 3 * 2^9: windows of 120 / 56 do not divide it
    k = (n & -n).bit_length() - 1
    exercise(uniform.UniformDgHierarchy(n, [3, 1], [2] * k, xin=0.0, xout=float(n), CDir=1000.0), n,
             "DG 3->1->agg (f_down/f_up 4x4, 2x2, f_tail)", pat)
    exercise(uniform.UniformDgHierarchy(n, [4, 2, 1], [2] * k, xin=0.0, xout=float(n), CDir=1000.0), n,
             "DG 4->2->1->agg (5x5, 3x3)", pat)
    rows = [{"pattern_resident": 0, "rows_window": 64, "rows_per_thread": 1},
            {"rows_window": 32, "rows_per_thread": 2}, {"rows_window": 64, "rows_per_thread": 3},
            {"pattern_resident": 1, "rows_per_thread": 0}, {"pattern_resident": 2}]
    exercise(uniform.UniformDgHierarchy(768, [8, 4, 2, 1], [2] * 8, xin=0.0, xout=768.0, CDir=1000.0), 768,
             "DG 8->4->2->1->agg (r_down/r_up 9x9)", rows)
    exercise(uniform.UniformCgHierarchy(n, [3, 1], [1], [2] * k, xin=0.0, xout=float(n), CDir=1000.0), n,
             "CG 3->1->DG 1->agg (point Jacobi, two-parent transfers)", pat)
    exercise(uniform.UniformCgHierarchy(768, [8, 4, 2, 1], [], [4] + [2] * 6, xin=0.0, xout=768.0, CDir=1000.0),
             768, "CG 8->4->2->1->agg (r_down/r_up on CG groups)", rows)

    # explicit per-element upload + device-side set-up (generic tier, k_galerkin, BCR coarse solve)
    import agglomerationmultigrid1d_b200 as aggmg
    from shapes import SHAPES, build_package
    for name in ("dg_heirarchy", "full_heirarchy", "bcr_dg_cg_n128"):
        for ds in (False, True):
            H, x0, b = build_package(device_setup=ds, **SHAPES[name])
            x, it, res, _ = aggmg.multigrid(H, x0, b, 100, 1e-10)
            print(f"[sanitize] {name} device_setup={ds}: {it} V-cycles", flush=True)
            H.device.close()
    print("[sanitize] done", flush=True)


if __name__ == "__main__":
    main()
