#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stages"; timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu --maxfail=30 -rf -k "hash or expert" > gpurun_out/pytest_stages.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_stages.log
echo "== probe"; timeout 300 python tools/grid_probe.py > gpurun_out/grid_probe.log 2>&1; echo "rc=$?"; grep -E "fwd_all|bwd_all" gpurun_out/grid_probe.log; grep "'level'" gpurun_out/grid_probe.log
