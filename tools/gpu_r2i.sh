#!/bin/bash
# Round-2 visit I: new tests, bench (full line kept), ncu launch list of the bench, ncu --set full of the routed path (CSV only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "== round2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_r2.log
echo "== bench";   timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/bench.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_field|k_adam|k_color|k_grad' -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/prof_routed_once.py > gpurun_out/prof_routed_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:'k_route_samples|k_blend|k_hashgrid_bwd_march_pts|k_bucket_plan|k_hashgrid_fwd|k_field_fwd_mma|k_field_bwd_mma' -c 40 -f -o /tmp/r02_routed python tools/prof_routed_once.py > gpurun_out/ncu_routed.log 2>&1; echo "ncu routed rc=$?"; tail -2 gpurun_out/ncu_routed.log
ncu -i /tmp/r02_routed.ncu-rep --page raw --csv > gpurun_out/r02_routed_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02_; du -sh gpurun_out
