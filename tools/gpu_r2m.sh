#!/bin/bash
# multi-GPU visit: sharded parity tests + bench at N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "== multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -rfs -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_multi.log
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_n$N.log
