#!/bin/bash
# ncu --set full of the level-0 legs of T in one of their forms (run on the B200 box via gpurun, AFTER the same
# command exited 0 without ncu).  usage: tools/ncu_legs.sh <tag> <kernel regex> <sweep spec>
#   tools/ncu_legs.sh r02q_pp 'f_(up|down)_pp' 4
#   tools/ncu_legs.sh r02q_dv 'f_(up|down)_dv' 4:leg_pipeline=0:dinv_registers=1
# Leaves the raw-metric and the source-page CSV under gpurun_out/ and removes the .ncu-rep (48 MB each; gpurun_out
# travels back only up to 64 MiB) unless KEEP_REP=1.
set -u
tag=$1; rx=$2; spec=$3
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$rx" -s 6 -c 2 -f -o gpurun_out/${tag}_full_L0 \
    python tools/sweep_dvrec.py T "$spec" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu $tag exit $?"
ncu -i gpurun_out/${tag}_full_L0.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_L0_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full_L0.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_ncu_source_sass.csv 2>/dev/null
[ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/${tag}_full_L0.ncu-rep
ls -la gpurun_out/${tag}_*
