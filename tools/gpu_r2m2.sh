#!/bin/bash
# Round-2 visit M2: warp-specialised fused backward (expert_bwd.cu): parity tests, A/B against the single-role kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc + round2 + render tests (new backward)"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py tests/test_gpu_render.py tests/test_gpu_train_step.py -q -m gpu --maxfail=20 -rf > gpurun_out/pytest_bwd_ws.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_bwd_ws.log
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
run ws1 ACN_BWD_WS=1
run ws0 ACN_BWD_WS=0
