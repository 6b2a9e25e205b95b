#!/bin/bash
# Round-2 visit R: standalone hash encode, per-lane paired 16-byte gathers vs 8-byte gathers (two-kernel forward for the A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_$name.log 2>&1; echo "== $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$name.log').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), {k:v['avg_ms'] for k,v in d['kernels'].items() if 'fwd' in k}, 'loss', d['e2e']['last_loss'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_$name.log').read()[-1500:])
PY
}
run enc_pair0 ACN_FUSED_FWD=0
run enc_pair1 ACN_FUSED_FWD=0 ACN_DEBUG_ENC_PAIR=1
