"""A/B of the pattern-resident modes on one workload: ms per V-cycle + check and level-0 leg times for option
pattern_resident = 0 / 1 / 2 (csrc/kernels_fused.cuh: PatOp table path, ParamOp constant-operand path).
Tuning builds of the library are selected with the AMG1D_LIB environment variable (one process per build).

  [AMG1D_LIB=build/libamg1d_c8.so] python tools/sweep_const.py [--workload T|C3|C4|C5|P8] [--modes 0,1,2]"""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="T")
    ap.add_argument("--log2n", type=int, default=0)
    ap.add_argument("--modes", default="0,1,2")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    log2n = args.log2n or bench.WORKLOADS[args.workload][0]
    n = 2 ** log2n
    torch.cuda.set_device(0)
    ts = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(ts)
    U = bench.build_hierarchy(args.workload, n)
    dev = U.upload(device=0, stream=ts.cuda_stream)
    upd = U.dof_updates_per_cycle()
    ref = None
    for mode in [int(m) for m in args.modes.split(",")]:
        dev.set_option("pattern_resident", mode)
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
        for l in range(min(3, len(U.levels) - 1)):
            for leg, nm in ((0, "down"), (1, "up")):
                t, cnt = dev.profile(l, leg)
                legs[f"L{l}_{nm}"] = round(t / max(cnt, 1), 4)
        dev.set_option("profile", 0)
        b_up = U.bytes_per_leg_fused(0, down=False)
        print(json.dumps({"lib": os.environ.get("AMG1D_LIB", "libamg1d.so"), "workload": args.workload, "n": n,
                          "pattern_resident": mode, "ms_per_cycle": ms, "dof_updates_per_s": upd / ms * 1e3,
                          "leg_ms": legs, "L0_up_GBps": b_up / legs["L0_up"] / 1e6,
                          "cycle_GBps": U.bytes_per_cycle_fused() / ms / 1e6,
                          "residual_identical_to_first": res == ref}), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
