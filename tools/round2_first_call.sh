#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out, in ONE gpurun call (about 10 GPU-minutes):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# 1. the validated suite (must stay green: the default kernels are SASS-identical, tests/test_abi.py checks that)
# 2. the gated tests of what has never run: bf16 MSDA, backward variants 7 / 8 / 10 / 11 / 12, the device transformer
#    against the reference's own DINOTransformer outputs
# 3. timings that decide which variant becomes the default: tools/bwd_variants.py, tools/bf16_msda.py
# 4. the TMA feed-rate microbenchmark behind the GEMM plan (docs/ROUND2_NOTES.md section 1)
# Results land in gpurun_out/r2_first_*.txt; each step has its own timeout so one hang cannot eat the call.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2_first_$name.txt" 2>&1; echo "rc=$? ($name)"; tail -n 5 "gpurun_out/r2_first_$name.txt"; }
run validated      420 python -m pytest tests -m gpu -x -q
SDB_RUN_UNVALIDATED=1 run bf16_tests     180 python -m pytest tests/test_msda_bf16_gpu.py -m gpu -q
SDB_RUN_UNVALIDATED=1 run bwd_exp_tests  300 python -m pytest tests/test_msda_bwd_experimental_gpu.py -m gpu -q
SDB_RUN_UNVALIDATED=1 run ref_golden_gpu 120 python -m pytest tests/test_dino_reference_golden.py -m gpu -q
run bwd_variants   120 python tools/bwd_variants.py
run bf16_timing    120 python tools/bf16_msda.py
run tma_rate       120 python tools/tma_rate.py
# 5. grad-weight split heuristic (8 k-blocks per split at least): parity first, then the step time with and without
SDB_GEMM_MIN_KBLOCKS_PER_SPLIT=8 run split_tests 240 python -m pytest tests/test_gemm_gpu.py tests/test_layernorm_gpu.py -m gpu -q
run bench_default  300 python bench.py --no-cpu-baseline --steps 20 --warmup 5
SDB_GEMM_MIN_KBLOCKS_PER_SPLIT=8 run bench_split8 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5
