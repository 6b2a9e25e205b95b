#!/bin/bash
# field-MLP iteration visit: tcgen05 tests, timing, timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_tc.log
echo "== prof"; timeout 300 python tools/prof_field.py > gpurun_out/prof_field.log 2>&1; echo "rc=$?"; cat gpurun_out/prof_field.log | tail -3
echo "== trace"; timeout 120 python tools/field_trace.py > gpurun_out/field_trace.log 2>&1; echo "rc=$?"; sed -n 1,60p gpurun_out/field_trace.log
