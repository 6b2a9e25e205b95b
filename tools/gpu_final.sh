#!/bin/bash
# Round-2 final visit (1 GPU): whole GPU suite, smoke, bench (both arms), ncu launch list of the bench, ncu --set full of the
# round-2 training-path kernels.  CSV summaries only travel back.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 -rfs > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/bench.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/bench_ref.log
echo "== configs"; timeout 900 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/bench_configs.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/r02_v6_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/prof_stages_r2.py > gpurun_out/prof_r2_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_' -f -o /tmp/r02_v6_stages python tools/prof_stages_r2.py > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"; tail -2 gpurun_out/ncu_stages.log
ncu -i /tmp/r02_v6_stages.ncu-rep --page raw --csv > gpurun_out/r02_v6_stages_raw.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/r02_v6_stages.ncu-rep k_expert_fwd 0 40 > gpurun_out/r02_v6_expert_fwd_stalls.txt 2>&1
timeout 300 python tools/prof_fused_bwd.py > gpurun_out/fused_bwd_ab.log 2>&1; tail -1 gpurun_out/fused_bwd_ab.log
ls -la gpurun_out | grep r02_v6; du -sh gpurun_out
