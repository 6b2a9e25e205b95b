#!/bin/bash
# 2-GPU visit: sharded-container parity + DP bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; cat gpurun_out/gpus.txt
echo "== multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -rfs > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_multi.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_n2.log
