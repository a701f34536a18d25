"""A/B sweep of the row-per-thread fused legs (csrc/kernels_rows.cuh): one upload, every compiled
(rows_window, rows_per_thread) variant, CUDA-event timing of whole V-cycles and of the level-0 legs.

  python tools/sweep_rows.py [--workload P8|CG8] [--log2n 23] [--steps 5]

Prints one JSON line per variant (ms per V-cycle + check, level-0 down / up leg ms, GB/s of the up leg)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    from agglomerationmultigrid1d_b200 import uniform
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="P8", choices=["P8", "CG8", "P6"])
    ap.add_argument("--log2n", type=int, default=23)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--pattern", action="store_true", help="sweep the variants with pattern-resident operators")
    args = ap.parse_args()
    n = 2 ** args.log2n
    torch.cuda.set_device(0)
    ts = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(ts)
    if args.workload == "P8":
        U = uniform.UniformDgHierarchy(n, [8, 4, 2, 1], [2] * args.log2n, xin=0.0, xout=float(n), CDir=1000.0)
    elif args.workload == "P6":
        U = uniform.UniformDgHierarchy(n, [6, 3, 1], [2] * args.log2n, xin=0.0, xout=float(n), CDir=1000.0)
    else:
        U = uniform.UniformCgHierarchy(n, [8, 4, 2, 1], [], [4] + [2] * (args.log2n - 2), xin=0.0, xout=float(n),
                                       CDir=1000.0)
    dev = U.upload(device=0, stream=ts.cuda_stream)
    dev.dev_fill_rhs_random(0)
    bytes_up = U.bytes_per_leg_fused(0, down=False)
    ref = None
    variants = [(w, r, 0) for w, r in ((64, 1), (32, 1), (32, 2), (64, 2), (32, 3), (64, 3))]
    if args.pattern:
        variants = [(64, 1, 0)] + [(w, r, 1) for w, r, _ in variants]
    for window, rpt, pat in variants:
        dev.set_option("rows_window", window)
        dev.set_option("rows_per_thread", rpt)
        dev.set_option("pattern_resident", pat)
        bytes_up = U.bytes_per_leg_fused(0, down=False)
        dev.dev_fill_rhs_random(0)
        for _ in range(3):
            dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            dev.dev_vcycle(with_residual_norm=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res = dev.dev_residual_norm()
        ref = res if ref is None else ref
        dev.set_option("profile", 1)
        for _ in range(args.steps):
            dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        legs = {}
        for leg, nm in ((0, "down"), (1, "up")):
            t, cnt = dev.profile(0, leg)
            legs[nm] = t / max(cnt, 1)
        dev.set_option("profile", 0)
        print(json.dumps({"workload": args.workload, "n": n, "rows_window": window, "rows_per_thread": rpt, "pattern_resident": pat,
                          "ms_per_cycle": ms, "L0_down_ms": legs["down"], "L0_up_ms": legs["up"],
                          "L0_up_GBps": bytes_up / legs["up"] / 1e6, "residual": res,
                          "residual_identical_to_first": res == ref}), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
