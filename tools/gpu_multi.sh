#!/bin/bash
# 2-GPU visit: sharded-container parity, sharded render/train timing, DP bench at N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "== multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -rfs > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_multi.log
echo "== sharded"; bash tools/gpu_sharded.sh $N
if [ "$2" == "bench" ]; then
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/bench_n$N.log
fi
