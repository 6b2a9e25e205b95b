#!/bin/bash
# One GPU-box visit: all parity tests, smoke, bench + its ncu launch list.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench";   timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3300 gpurun_out/bench.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -c 900 gpurun_out/bench_ref.log
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
fi
