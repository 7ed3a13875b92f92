#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -5
timeout 300 python tools/time_ffn_dgrad.py 2>&1 | tee gpurun_out/r2d3_ffn_dgrad.json | tail -3
timeout 300 python tools/time_linear1.py 2>&1 | tee gpurun_out/r2d3_linear1.json | tail -3
