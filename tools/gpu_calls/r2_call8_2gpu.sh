#!/usr/bin/env bash
# round 2, GPU call 8 (2 GPUs): residual+LayerNorm fusion, overlapped gradient exchange (fixed cut), N=1/2 benches
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c8_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 4 "gpurun_out/r2c8_$name.txt" | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run ln_tests 300 python -m pytest tests/test_layernorm_gpu.py tests/test_dino_gpu.py tests/test_dino_reference_golden.py -m gpu -q -s
run check_overlap 400 $TR --master-port 29511 tools/check_overlap.py
run bench_n2_overlap 500 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5
SDB_OVERLAP=0 run bench_n2_single 500 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5
run bench_n1 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run suite 900 python -m pytest tests -m gpu -q
