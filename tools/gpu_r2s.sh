#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "fused_expert_backward" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_x.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_x.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',round(d['ms_per_step'],3),'bwd',d['kernels']['acn_render_expert_bwd']['avg_ms'],'fwd',d['kernels']['acn_render_expert_fwd']['avg_ms'])
"
