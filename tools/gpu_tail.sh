#!/bin/bash
# Loss epilogue + optimizer tail: parity tests, their timing against PyTorch, then the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== train_step tests"; timeout 600 python -m pytest tests/test_gpu_train_step.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_train_step.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_train_step.log
echo "== tail timing"; timeout 300 python tools/bench_tail.py > gpurun_out/bench_tail.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_tail.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3300 gpurun_out/bench.log
