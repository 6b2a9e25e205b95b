#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc tests (ws)"; ACN_BWD_WS=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "fused_expert_backward" > gpurun_out/pytest_ws.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_ws.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_$name.log 2>&1; echo "== $name rc=$?"
  tail -1 gpurun_out/bench_$name.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/step',round(d['ms_per_step'],3),'bwd',d['kernels']['acn_render_expert_bwd']['avg_ms'],'loss',d['e2e']['last_loss'])
" || tail -20 gpurun_out/bench_$name.log
}
run ws8 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8
run ws8_c4 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8 ACN_BWD_CHAIN_LEVELS=4
run ws8_c6 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8 ACN_BWD_CHAIN_LEVELS=6
run ws8_c8 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8 ACN_BWD_CHAIN_LEVELS=8
for c in 4 8; do echo "== tc tests (ws8 c$c)"; ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8 ACN_BWD_CHAIN_LEVELS=$c timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "fused_expert_backward" > gpurun_out/pytest_ws8c$c.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_ws8c$c.log; done
