#!/bin/bash
# full ncu capture of the constant-operand level-0 legs (pattern_resident = 2) of workload T
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:f_(up|down)_c<\(int\)4' -s 6 -c 2 -f -o gpurun_out/${1:-r01e}_full_L0_const \
    python tools/sweep_const.py --modes 2 --steps 2 > gpurun_out/ncu_const.log 2>&1
echo "full capture exit $?"
