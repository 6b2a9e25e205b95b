#!/bin/bash
# 2-GPU visit: whole GPU suite incl. the two-GPU cases, then bench.py --gpus 2 as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 -rfs > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu2.log
bash tools/gpu_bench_n.sh 2
