#!/bin/bash
# Round-2 visit K: fused forward A/B (paired gathers, producer warps), new tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fused fwd + bg tests"; timeout 600 python -m pytest tests/test_gpu_fused_fwd.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_fused_fwd.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_fused_fwd.log
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_$name.log 2>&1; echo "== $name rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$name.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'render ms', round(d['render']['ms_per_batch'],3), 'fwd', d['kernels'].get('acn_render_expert_fwd',{}).get('avg_ms'))
PY
}
run p0_np16 ACN_FWD_PAIR=0 ACN_FWD_NP=16
run p0_np20 ACN_FWD_PAIR=0 ACN_FWD_NP=20
run p0_np24 ACN_FWD_PAIR=0 ACN_FWD_NP=24
run p1_np16 ACN_FWD_PAIR=1 ACN_FWD_NP=16
run p1_np24 ACN_FWD_PAIR=1 ACN_FWD_NP=24
