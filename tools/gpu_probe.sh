#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== umma probe"; timeout 600 python tools/umma_probe.py > gpurun_out/umma_probe.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/umma_probe.log
echo "== grid probe"; timeout 300 python tools/grid_probe.py > gpurun_out/grid_probe.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/grid_probe.log
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"
echo "== ncu full (hashgrid fwd/bwd, field fwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hashgrid|k_field_fwd_tc' -s 12 -c 3 -o gpurun_out/r01_prof python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
