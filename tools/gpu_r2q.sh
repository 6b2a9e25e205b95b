#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/red_probe.py gpurun_out/r02_red_probe.jsonl > gpurun_out/red_probe.log 2>&1; echo "rc=$?"; tail -50 gpurun_out/red_probe.log
