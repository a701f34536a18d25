#!/usr/bin/env python
"""A/B of option recompute_dinv on one workload: per-leg device times (CUDA events, option profile) and ms per
V-cycle + check (graph replay) for several threshold values, one upload.  JSON lines on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    workload = sys.argv[1] if len(sys.argv) > 1 else "T"
    # a value is  <recompute_dinv>[:key=val[:key=val]]  (extra options for that variant, e.g. 4:leg_pipeline=0:dinv_registers=1)
    specs = (sys.argv[2] if len(sys.argv) > 2 else "0,2,3,4").split(",")
    values = [int(v.split(":")[0]) for v in specs]
    extras = [dict(kv.split("=") for kv in v.split(":")[1:]) for v in specs]
    log2n = int(sys.argv[3]) if len(sys.argv) > 3 else bench.WORKLOADS[workload][0]    # optional size override
    U = bench.build_hierarchy(workload, 2 ** log2n)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    dev = U.upload(stream=ts.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for v, extra in zip(values, extras):
        dev.set_option("dinv_registers", 0)
        dev.set_option("leg_pipeline", 1)
        dev.set_option("leg_pipeline_min", 500000)                # the library default
        dev.set_option("recompute_dinv_max", 4)
        for k, val in extra.items():
            dev.set_option(k, int(val))
        dev.set_option("recompute_dinv", v)
        dev.dev_fill_rhs_random(0)
        for _ in range(3):
            dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        ev0.record()
        for _ in range(10):
            dev.dev_vcycle(with_residual_norm=True)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        res = dev.dev_residual_norm()
        ref = res if ref is None else ref
        dev.set_option("profile", 1)
        for _ in range(5):
            dev.dev_vcycle(with_residual_norm=True)
        dev.synchronize()
        legs = {}
        for l in range(min(5, len(U.levels) - 1)):
            for leg, nm in ((0, "down"), (1, "up")):
                t, c = dev.profile(l, leg)
                legs[f"L{l}_{nm}"] = round(t / max(c, 1), 4)
        dev.set_option("profile", 0)
        print(json.dumps({"workload": workload, "log2n": log2n, "recompute_dinv": v, "extra": extra, "ms_per_cycle": ms, "legs": legs,
                          "recomputing_levels": [l for l in range(len(U.levels) - 1) if dev.info(f"dinv_recompute:{l}") == 1][:8],
                          "pivots": [dev.info(f"dinv_pivots:{l}") for l in range(4)],
                          "pipelined_levels": [l for l in range(len(U.levels) - 1) if dev.info(f"leg_pipeline:{l}") == 1][:8],
                          "bytes_per_cycle": U.bytes_per_cycle_fused(), "residual_bit_identical": res == ref}), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
