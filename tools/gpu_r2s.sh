#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_$name.log 2>&1; echo "== $name rc=$?"
  tail -1 gpurun_out/bench_$name.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',round(d['ms_per_step'],3),'fwd',d['kernels']['acn_render_expert_fwd']['avg_ms'],'render',round(d['render']['ms_per_batch'],3))
" || tail -20 gpurun_out/bench_$name.log
}
run np12 A=1
