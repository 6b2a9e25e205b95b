#!/bin/bash
# ncu --set full of the two fused MLP kernels at 2^21 samples (second launch of each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/prof_field.py --small > gpurun_out/prof_field_small.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_(fwd|bwd)_mma' -s 2 -c 2 -f -o gpurun_out/field_prof python tools/prof_field.py --small > gpurun_out/ncu_field.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_field.log
ncu -i gpurun_out/field_prof.ncu-rep --page raw --csv > gpurun_out/field_prof_raw.csv 2>/dev/null
