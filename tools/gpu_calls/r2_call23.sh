#!/usr/bin/env bash
# round 2, GPU call 23: padded-batch step test, mask-skip equivalence, library ReLU epilogue for linear1, benches
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c23_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 3 "gpurun_out/r2c23_$name.txt" | cut -c1-400; }
run dino_tests  600 python -m pytest tests/test_dino_gpu.py tests/test_gemm_gpu.py tests/test_ssod_device_gpu.py tests/test_attention_gpu.py -m gpu -q -s
run bench_sup   400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
SDB_SKIP_EMPTY_MASK=0 run bench_sup_masked 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run suite       900 python -m pytest tests -m gpu -q
