#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in "a1f1:SDB_ATTN_BLOCK=1 SDB_FOLD_RESIDUAL=1" "a1f0:SDB_ATTN_BLOCK=1 SDB_FOLD_RESIDUAL=0" "a0f1:SDB_ATTN_BLOCK=0 SDB_FOLD_RESIDUAL=1" "a0f0:SDB_ATTN_BLOCK=0 SDB_FOLD_RESIDUAL=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d10_sup_$name.json 2> gpurun_out/r2d10_sup_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2d10_sup_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["e2e"]["last_loss"], d["gpu_launches_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2d10_sup_$name.err").read()[-1500:])
P
done
