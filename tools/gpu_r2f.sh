#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stages+tc"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_round2.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_tc.log
echo "== fused bwd timing"; timeout 300 python tools/prof_fused_bwd.py > gpurun_out/fused_bwd.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/fused_bwd.log
