#!/bin/bash
# ncu evidence for profiles/ (run on the B200 box via gpurun, AFTER the same bench command exited 0 without ncu):
#   1. launch list of the default bench command (gpu__time_duration.sum per launch)
#   2. full capture of the level-0 legs f_down<4,..> / f_up<4,..> with pattern-resident operators
# Numbers printed by a run under ncu are never bench values.
set -u
tag=${1:-r01d}
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 420 --csv \
    --log-file gpurun_out/${tag}_launches_T.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:f_(up|down)<\(int\)4' -s 6 -c 2 -f -o gpurun_out/${tag}_full_L0_pattern \
    python bench.py --steps 2 --warmup 3 --no-cpu --opt pattern_resident=1 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
