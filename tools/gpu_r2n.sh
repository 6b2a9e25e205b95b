#!/bin/bash
# multi-GPU bench only: tools/gpu_r2n.sh N [extra bench flags]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
echo "== bench N=$N $@"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 "$@" > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; grep -o '"expert_sharded".*' gpurun_out/bench_n$N.log | cut -c1-2500; grep -o '"ms_per_step": [0-9.]*, "higher' gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.log | cut -c1-400 | grep -v '^{'
