#!/usr/bin/env bash
# round 2, GPU call 16 (2 GPUs): gradient exchange fused with the optimizer over NVLink multicast
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c16_$name.txt" 2>&1; echo "rc=$? ($name)"; grep '^{' "gpurun_out/r2c16_$name.txt" | cut -c1-900 || tail -n 5 "gpurun_out/r2c16_$name.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run check_exchange 120 $TR --master-port 29517 tools/check_exchange.py
run check_overlap 300 $TR --master-port 29511 tools/check_overlap.py
run bench_n2_peer 400 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5
SDB_EXCHANGE=nccl run bench_n2_nccl 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5
run bench_n1 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
