#!/bin/bash
# pipelined persistent legs (leg_pipeline_min = 0) against the one-window-per-CTA legs over problem sizes 2^18 .. 2^24 (T shape):
# the measurement behind the default of option leg_pipeline_min (profiles/r03a_sweep_sizes.jsonl)
for n in 18 20 21 22 23 24; do timeout 120 python tools/sweep_dvrec.py T "4:leg_pipeline_min=0,4:leg_pipeline_min=99999999,4:leg_pipeline_min=99999999:dinv_registers=1,4:leg_pipeline_min=0,4:leg_pipeline_min=99999999,4:leg_pipeline_min=99999999:dinv_registers=1" $n >> gpurun_out/r03a_sweep_sizes.jsonl 2>> gpurun_out/sweep.err; done
python - <<PY
import json
for line in open("gpurun_out/r03a_sweep_sizes.jsonl"):
    d=json.loads(line); print(d["log2n"], d["extra"], round(d["ms_per_cycle"],4), d["legs"]["L0_down"], d["legs"]["L0_up"], d["pipelined_levels"], d["residual_bit_identical"])
PY
tail -3 gpurun_out/sweep.err
