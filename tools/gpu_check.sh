#!/bin/bash
# One GPU-box visit: parity tests (tcgen05 tests isolated under their own timeout so a bad
# descriptor cannot hang the box), smoke, a short bench.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stages";  timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_stages.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_stages.log
echo "== tc";      timeout 180 python -m pytest tests/test_gpu_tc.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_tc.log
echo "== render";  timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_render.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_render.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench";   timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench.log
