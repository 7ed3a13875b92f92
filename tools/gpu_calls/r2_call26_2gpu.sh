#!/usr/bin/env bash
# round 2, GPU call 26 (2 GPUs): teacher-student step with frozen state at N=1 / N=2; tile tests with the L1 prefetch
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c26_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2c26_$name.txt" || tail -n 4 "gpurun_out/r2c26_$name.txt") | cut -c1-300; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run tile_tests 300 python -m pytest tests/test_msda_tile_gpu.py tests/test_msda_fused_gpu.py -m gpu -q
run bench_ssod_n1 900 python bench.py --workload ssod --steps 5 --warmup 5 --no-cpu-baseline
run bench_ssod_n2 900 $TR --master-port 29513 bench.py --gpus 2 --workload ssod --steps 5 --warmup 5
