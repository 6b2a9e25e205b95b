#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
: > gpurun_out/bench_sharded.log
for W in 1 $N; do
  for MODE in "" "--train" "--peer" "--train --peer"; do
    if [ "$W" == "1" ] && [[ "$MODE" == *peer* ]]; then continue; fi
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29544 tools/bench_sharded.py $MODE 2>&1 | grep -E '^\{|Error|error' >> gpurun_out/bench_sharded.log
  done
done
cat gpurun_out/bench_sharded.log
