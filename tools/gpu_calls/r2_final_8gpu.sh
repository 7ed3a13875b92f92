#!/usr/bin/env bash
# round 2, final 8-GPU call: configs[1] at N=8 and N=4 (fused peer exchange), configs[2] per-GPU batch at N=8
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2f8_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2f8_$name.txt" || tail -n 3 "gpurun_out/r2f8_$name.txt") | cut -c1-260; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run bench_sup_n8 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5
run bench_sup_n4 300 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5
run bench_ssod_n8 600 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --workload ssod --steps 5 --warmup 8
