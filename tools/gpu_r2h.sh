#!/bin/bash
# Round-2 visit H: full suite, smoke, bench (N=1), reference arm, ncu launch list of the bench, ncu --set full of the routed path.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 -rf > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench";   timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/bench.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/bench_ref.log
echo "== configs"; timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/bench_configs.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_field|k_adam|k_color|k_grad' -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/prof_routed_once.py > gpurun_out/prof_routed_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_route_samples|k_blend|k_hashgrid_bwd_march_pts|k_bucket_plan|k_hashgrid_fwd|k_field_fwd_mma|k_field_bwd_mma' -c 40 -f -o gpurun_out/r02_routed python tools/prof_routed_once.py > gpurun_out/ncu_routed.log 2>&1; echo "ncu routed rc=$?"; tail -2 gpurun_out/ncu_routed.log
ncu -i gpurun_out/r02_routed.ncu-rep --page raw --csv > gpurun_out/r02_routed_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02_
