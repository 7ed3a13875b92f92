#!/usr/bin/env bash
# round 2, GPU call 9: bf16 5-scale step (test + bench), fresh step profile, launch list of the supervised step
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c9_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 4 "gpurun_out/r2c9_$name.txt" | cut -c1-900; }
run bf16_test   400 python -m pytest tests/test_dino_gpu.py -m gpu -q -s -k "bf16 or tf32"
run bench_sup5  600 python bench.py --workload sup5 --steps 10 --warmup 3
run profile     300 python tools/profile_step.py
run launches    600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c9_launches.csv python bench.py --ncu
