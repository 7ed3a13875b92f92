#!/usr/bin/env bash
# round 2, GPU call 7 (2 GPUs): overlapped gradient exchange -- equality check, then N=2 bench with and without overlap
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c7_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 4 "gpurun_out/r2c7_$name.txt" | cut -c1-600; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run check_overlap 400 $TR --master-port 29511 tools/check_overlap.py
run bench_n2_overlap 500 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5
SDB_OVERLAP=0 run bench_n2_single 500 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5
run bench_n1 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
