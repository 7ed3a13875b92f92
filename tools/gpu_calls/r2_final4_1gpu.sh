#!/usr/bin/env bash
# last check of HEAD: smoke, full GPU suite, configs[1] bench line
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_sup.txt 2>&1; grep '^{' gpurun_out/r2i_bench_sup.txt | cut -c1-330
