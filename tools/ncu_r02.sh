#!/bin/bash
# ncu evidence of round 2 (run on the B200 box via gpurun, AFTER the same bench command exited 0 without ncu):
#   1. launch list of the default bench command (gpu__time_duration.sum per launch)
#   2. full capture of the level-0 legs f_down<4,..> / f_up<4,..> of T with the smoother inverse recomputed in
#      registers (the default) - DRAM traffic per launch goes to profiles/traffic.json
# Numbers printed by a run under ncu are never bench values.
set -u
tag=${1:-r02}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 420 --csv \
    --log-file gpurun_out/${tag}_launches_T.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-pattern > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:f_(up|down)(_pp|_dv)?<\(int\)4' -s 6 -c 2 -f -o gpurun_out/${tag}_full_L0 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-pattern > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/${tag}_full_L0.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_L0_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full_L0.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_ncu_source_sass.csv 2>/dev/null
[ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/${tag}_full_L0.ncu-rep    # 48 MB each; gpurun_out travels back only up to 64 MiB
ls -la gpurun_out/${tag}_*
