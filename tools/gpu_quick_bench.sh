#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_fwd.py -q -m gpu -k "queued" 2>&1 | tail -1
timeout 600 python bench.py --no-extras > gpurun_out/bench_quick.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'], d['clocks'])
" || tail -30 gpurun_out/bench_quick.log
