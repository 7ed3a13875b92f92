#!/usr/bin/env bash
# round 2, GPU call 10: fused decoder self-attention (tests, step bench), bf16 5-scale test
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c10_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 4 "gpurun_out/r2c10_$name.txt" | cut -c1-900; }
run attn_tests  300 python -m pytest tests/test_attention_gpu.py -m gpu -q -s
run dino_tests  600 python -m pytest tests/test_dino_gpu.py tests/test_dino_reference_golden.py -m gpu -q -s
run bench_sup   400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
SDB_ATTENTION=eager run bench_sup_eager_attn 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
