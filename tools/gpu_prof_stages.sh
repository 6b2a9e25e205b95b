#!/bin/bash
# ncu --set full over ONE launch of every stage kernel (second repetition of tools/prof_stages.py), summaries as CSV.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L2N=${1:-18}
timeout 300 python tools/prof_stages.py $L2N > gpurun_out/prof_stages_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'^k_' -f -o gpurun_out/r01_stages python tools/prof_stages.py $L2N > gpurun_out/ncu_stages.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_stages.log
ncu -i gpurun_out/r01_stages.ncu-rep --page raw --csv > gpurun_out/r01_stages_raw.csv 2>/dev/null
ls -la gpurun_out/
SZ=$(stat -c %s gpurun_out/r01_stages.ncu-rep); if [ "$SZ" -gt 40000000 ]; then rm gpurun_out/r01_stages.ncu-rep; fi
