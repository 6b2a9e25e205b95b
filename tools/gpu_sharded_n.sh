#!/bin/bash
# expert-sharded timings at N GPUs only (NCCL and peer-memory exchange) + DP bench at N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
: > gpurun_out/bench_sharded_n$N.log
for MODE in "" "--train" "--peer" "--train --peer"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/bench_sharded.py $MODE 2>&1 | grep -E '^\{|Error|error' >> gpurun_out/bench_sharded_n$N.log
done
cat gpurun_out/bench_sharded_n$N.log
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; tail -c 400 gpurun_out/bench_n$N.log | head -c 400; grep -o '"value": [0-9.]*, "unit": "rays/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench_n$N.log
