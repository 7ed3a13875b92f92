#!/usr/bin/env bash
# final code on 2 GPUs: exchange tests + configs[1] bench line
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2d13_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2d13_$name.txt" || tail -n 4 "gpurun_out/r2d13_$name.txt") | cut -c1-300; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run exchange_tests 400 python -m pytest tests/test_exchange_gpu.py tests/test_engine_gpu.py -m gpu -q
run bench_sup_n2 600 $TR --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5
