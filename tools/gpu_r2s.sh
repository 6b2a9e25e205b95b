#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_render.py tests/test_gpu_fused_fwd.py tests/test_gpu_tc.py -q -m gpu 2>&1 | tail -2
timeout 600 python tools/bench_configs.py 2>&1 | tail -3 | cut -c1-400
timeout 600 python tools/prof_frame.py 2>&1 | head -6
