#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (+ --graph timing of the sharded step), both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --graph > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])
print(json.dumps(d.get('expert_sharded'))[-1400:])
print(json.dumps(d.get('frame'))[-200:])
" || tail -30 gpurun_out/bench_n$N.log
echo "== reference arm N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "rc=$?"; tail -c 200 gpurun_out/bench_ref_n$N.log
