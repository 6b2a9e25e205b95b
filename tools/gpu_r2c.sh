#!/bin/bash
# Round-2 visit C: new parity tests, fused backward timing, ncu --set full of the fused backward.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== round2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_tc.py tests/test_gpu_train_step.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_r2.log
echo "== fused bwd timing"; timeout 300 python tools/prof_fused_bwd.py > gpurun_out/fused_bwd.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/fused_bwd.log
echo "== bench";   timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/bench.log
timeout 300 python tools/prof_fused_bwd.py 18 --once > gpurun_out/fused_once_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_field_bwd_mma|k_hashgrid_bwd_march' -c 3 -o gpurun_out/r02_fused_bwd python tools/prof_fused_bwd.py 18 --once > gpurun_out/ncu_fused.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log
