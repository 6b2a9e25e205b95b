#!/bin/bash
# ncu --set full of the warp-specialised fused backward (k_expert_bwd, 8 scatter warps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8
timeout 300 python tools/prof_fused_bwd.py 18 --once > gpurun_out/fused_once_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/fused_once_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_expert_bwd' -c 1 -f -o /tmp/r02_ws_bwd python tools/prof_fused_bwd.py 18 --once > gpurun_out/ncu_ws_bwd.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ws_bwd.log
ncu -i /tmp/r02_ws_bwd.ncu-rep --page raw --csv > gpurun_out/r02_ws8_bwd_raw.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/r02_ws_bwd.ncu-rep k_expert_bwd 0 60 > gpurun_out/r02_ws8_bwd_stalls.txt 2>&1
ls -la gpurun_out | grep r02_ws8
