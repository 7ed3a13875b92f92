#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/time_ffn_dgrad.py 2>&1 | tee gpurun_out/r2d6_ffn_dgrad.json | tail -3
for v in "block2:SDB_GEMM_RELU_GRAD_ROUND=2" "noblock:SDB_FFN_BLOCK=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d6_sup_$name.json 2> gpurun_out/r2d6_sup_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2d6_sup_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["e2e"]["last_loss"], d["gpu_launches_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2d6_sup_$name.err").read()[-1500:])
P
done
