#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_fused_gpu.py -x -q -k "encoder_attention_block" 2>&1 | tail -12
for v in "blk:SDB_ATTN_BLOCK=1" "noblk:SDB_ATTN_BLOCK=0" "blk2:SDB_ATTN_BLOCK=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d9_sup_$name.json 2> gpurun_out/r2d9_sup_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2d9_sup_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["e2e"]["last_loss"], d["gpu_launches_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2d9_sup_$name.err").read()[-1500:])
P
done
