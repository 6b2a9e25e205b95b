#!/bin/bash
# Round-2 visit J: fused forward (acn_render_expert_fwd): parity tests, whole GPU suite, bench A/B (fused / staged on and off).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fused fwd tests"; timeout 600 python -m pytest tests/test_gpu_fused_fwd.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_fused_fwd.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_fused_fwd.log
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -rf > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for v in "1 1" "1 0" "0 0"; do set -- $v
  echo "== bench fused=$1 staged=$2"; ACN_FUSED_FWD=$1 ACN_STAGE_COARSE=$2 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f$1_s$2.log 2>&1; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f$1_s$2.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'render ms', round(d['render']['ms_per_batch'],3), 'frame', d.get('frame',{}).get('ms_per_frame'))
print({k:v['avg_ms'] for k,v in d['kernels'].items()})
PY
done
