#!/usr/bin/env bash
# round 2, GPU call 6: vectorised warm-up phase, TF32 replay test, SSOD profiles
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c6_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 6 "gpurun_out/r2c6_$name.txt"; }
run suite        900 python -m pytest tests -m gpu -q -s
run bench_ssod   900 python bench.py --workload ssod --steps 5 --warmup 5 --no-cpu-baseline
run prof_ssod_w  300 python tools/profile_ssod.py 0
run prof_ssod_h  300 python tools/profile_ssod.py 60000
run bench_sup    400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
