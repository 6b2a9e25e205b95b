#!/bin/bash
# Round-2 visit A: fused expert backward (parity + timing), L2 peaks, full GPU suite, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tc"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_tc.log
echo "== fused bwd timing"; timeout 300 python tools/prof_fused_bwd.py > gpurun_out/fused_bwd.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/fused_bwd.log
echo "== l2 peaks"; timeout 300 python tools/l2_peak.py gpurun_out/l2_peaks.json > gpurun_out/l2_peaks.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/l2_peaks.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=30 -rf --deselect tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench";   timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3300 gpurun_out/bench.log
