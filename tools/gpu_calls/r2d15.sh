#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for mk in 0 4 8 16; do
  SDB_GEMM_MIN_KBLOCKS_PER_SPLIT=$mk timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "grad_weight or linear_layer_matches" 2>&1 | tail -1
  SDB_GEMM_MIN_KBLOCKS_PER_SPLIT=$mk timeout 300 python tools/profile_step.py > gpurun_out/r2d15_profile_mk$mk.txt 2>&1
  echo "mk=$mk"; grep -E "GPU kernel time|gemm_tf32_kernel<true, true, false>|Memset|FillFunctor" gpurun_out/r2d15_profile_mk$mk.txt | cut -c1-120
done
