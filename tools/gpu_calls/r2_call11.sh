#!/usr/bin/env bash
# round 2, GPU call 11: tile backward v5 (pixel runs per lane group, FFMA2, byte-offset visit words)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c11_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 6 "gpurun_out/r2c11_$name.txt" | cut -c1-600; }
run tile_tests   300 python -m pytest tests/test_msda_tile_gpu.py tests/test_msda_fused_gpu.py -m gpu -x -q
run variants     200 python tools/bwd_variants.py
run bench_msda   300 python bench.py --workload msda --steps 30 --warmup 5 --no-cpu-baseline
run ncu_tile     400 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_tile -s 2 -c 1 -o gpurun_out/r2c11_ncu_tile python tools/bwd_variants.py
