#!/bin/bash
# Routing / bucketing kernels: parity tests, then the per-kernel split of the 1080p frame and configs 3-5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_render.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_route.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_route.log
echo "== frame"; timeout 600 python tools/prof_frame.py > gpurun_out/prof_frame.log 2>&1; echo "rc=$?"; head -14 gpurun_out/prof_frame.log
echo "== configs"; timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_configs.log
