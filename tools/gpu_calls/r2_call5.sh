#!/usr/bin/env bash
# round 2, GPU call 5: tile backward v4, SSOD device kernels (NMS + filter, GMM), full suite, benches
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2c5_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 6 "gpurun_out/r2c5_$name.txt"; }
run tile_tests   300 python -m pytest tests/test_msda_tile_gpu.py tests/test_msda_fused_gpu.py -m gpu -x -q
run ssod_dev     300 python -m pytest tests/test_ssod_device_gpu.py -m gpu -q
run variants     200 python tools/bwd_variants.py
run suite        900 python -m pytest tests -m gpu -q -s
run bench_msda   300 python bench.py --workload msda --steps 30 --warmup 5 --no-cpu-baseline
run bench_sup    400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run ncu_tile     400 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_tile -s 2 -c 1 -o gpurun_out/r2c5_ncu_tile python tools/bwd_variants.py
run bench_ssod   900 python bench.py --workload ssod --steps 5 --warmup 3 --no-cpu-baseline
