#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== configs"; timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/bench_configs.log
echo "== bench";   timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/bench.log
