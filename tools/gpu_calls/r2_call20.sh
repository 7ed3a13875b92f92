#!/usr/bin/env bash
# round 2, GPU call 20: smoke, full suite, default bench line, linear-policy comparison, clean launch list of one step
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c20_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 3 "gpurun_out/r2c20_$name.txt" | cut -c1-500; }
run smoke       300 python __graft_entry__.py --smoke
run suite       900 python -m pytest tests -m gpu -q
run bench_sup   600 python bench.py --steps 20 --warmup 5
SDB_LINEAR=cublas run bench_cublas 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
SDB_LINEAR=tcgen05 run bench_tcgen05 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run launches    600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c20_launches.csv python bench.py --ncu
