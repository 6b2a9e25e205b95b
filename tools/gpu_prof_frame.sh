#!/bin/bash
# ncu --set full over the hash-encode launches of one frame chunk in both warp mappings (tools/prof_frame_mapping.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/prof_frame_mapping.py > gpurun_out/prof_frame_mapping.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_hashgrid_fwd' -f -o gpurun_out/r01_frame python tools/prof_frame_mapping.py > gpurun_out/ncu_frame.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_frame_mapping.log | tail -3; tail -2 gpurun_out/ncu_frame.log
ncu -i gpurun_out/r01_frame.ncu-rep --page raw --csv > gpurun_out/r01_frame_raw.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/r01_frame.ncu-rep); if [ "$SZ" -gt 40000000 ]; then rm gpurun_out/r01_frame.ncu-rep; fi
ls -la gpurun_out/ | grep frame
