#!/usr/bin/env bash
# round 2, GPU call 17 (8 GPUs): fused exchange at world size 8, configs[1] step at N=8
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c17_$name.txt" 2>&1; echo "rc=$? ($name)"; grep '^{' "gpurun_out/r2c17_$name.txt" | cut -c1-700 || tail -n 5 "gpurun_out/r2c17_$name.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run check_exchange 120 $TR --master-port 29517 tools/check_exchange.py
run bench_n8_peer 300 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5
SDB_EXCHANGE=nccl run bench_n8_nccl 300 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5
