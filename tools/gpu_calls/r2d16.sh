#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in b1a:1 b0a:0 b1b:1 b0b:0 b1c:1 b0c:0; do
  name=${v%%:*}; flag=${v#*:}
  SDB_CUDNN_BENCHMARK=$flag timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d16_sup_$name.json 2> gpurun_out/r2d16_sup_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2d16_sup_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2d16_sup_$name.err").read()[-1500:])
P
done
