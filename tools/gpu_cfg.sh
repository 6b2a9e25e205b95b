#!/bin/bash
# BASELINE configs 3-5 on one GPU + the per-kernel split of the 1080p frame, after the tests touched by the change.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_render.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_cfg.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_cfg.log
echo "== configs"; timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_configs.log
echo "== frame"; timeout 600 python tools/prof_frame.py > gpurun_out/prof_frame.log 2>&1; echo "rc=$?"; head -60 gpurun_out/prof_frame.log
