#!/usr/bin/env bash
# round 2, GPU call 25 (2 GPUs): teacher-student step at N=2, supervised step at N=2 with the current code, exchange tests
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c25_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2c25_$name.txt" || tail -n 4 "gpurun_out/r2c25_$name.txt") | cut -c1-600; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run exchange_tests 400 python -m pytest tests/test_exchange_gpu.py -m gpu -q
run bench_sup_n2 400 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5
run bench_ssod_n2 900 $TR --master-port 29513 bench.py --gpus 2 --workload ssod --steps 5 --warmup 5
run bench_ssod_n1 900 python bench.py --workload ssod --steps 5 --warmup 5 --no-cpu-baseline
