#!/bin/bash
# Round-2 visit L (2 GPUs): whole GPU suite incl. the two-GPU cases (C-ABI communicator), then the bench at N=1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 -rfs > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 400 gpurun_out/bench.log
