#!/bin/bash
# Round-2 visit P: warp-specialised backward A/B (batched REDs; chain-only timing)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu --maxfail=20 -rf > gpurun_out/pytest_bwd_ws.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_bwd_ws.log
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_$name.log 2>&1; echo "== $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$name.log').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'bwd', d['kernels'].get('acn_render_expert_bwd',{}).get('avg_ms'), 'fwd', d['kernels'].get('acn_render_expert_fwd',{}).get('avg_ms'), 'loss', d['e2e']['last_loss'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_$name.log').read()[-1500:])
PY
}
run ws16 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=16
run ws8 ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8
run ws8_nored ACN_BWD_WS=1 ACN_BWD_SCATTER_WARPS=8 ACN_DEBUG_NO_RED=1
