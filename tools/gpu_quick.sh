#!/bin/bash
# Quick iteration visit: the tests touched by the current change + bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_tc.log
echo "== render"; timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_render.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_render.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 2600 gpurun_out/bench.log
