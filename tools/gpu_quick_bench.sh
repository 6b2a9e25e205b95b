#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_quick.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'])
" || tail -30 gpurun_out/bench_quick.log
