#!/usr/bin/env bash
# round 2, HEAD at N=8: configs[1] bench line (fused peer exchange)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h8_bench_sup_n8.txt 2>&1
echo "rc=$?"; (grep '^{' gpurun_out/r2h8_bench_sup_n8.txt || tail -n 5 gpurun_out/r2h8_bench_sup_n8.txt) | cut -c1-300
