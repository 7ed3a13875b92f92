#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2d8_suite_tail.txt
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d8_sup.json 2> gpurun_out/r2d8_sup.err; cut -c1-400 gpurun_out/r2d8_sup.json
timeout 300 python tools/profile_step.py > gpurun_out/r2d8_profile.txt 2>&1; head -5 gpurun_out/r2d8_profile.txt
