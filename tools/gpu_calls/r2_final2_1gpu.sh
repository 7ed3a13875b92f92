#!/usr/bin/env bash
# round 2, final single-GPU validation: smoke, full suite, every bench workload, reference arm, launch list, ncu of the
# dominant kernel.  Outputs under gpurun_out/r2g_*.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2g_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2g_$name.txt" || tail -n 3 "gpurun_out/r2g_$name.txt") | cut -c1-260; }
run smoke       300 python __graft_entry__.py --smoke
run suite       900 python -m pytest tests -m gpu -q
run bench_sup   600 python bench.py --steps 20 --warmup 5
run bench_ref   600 python bench.py --impl reference --steps 2 --warmup 1
run bench_msda  400 python bench.py --workload msda --steps 100 --warmup 10
run bench_sup5  600 python bench.py --workload sup5 --steps 10 --warmup 3
run bench_ssod  900 python bench.py --workload ssod --steps 8 --warmup 8 --no-cpu-baseline
run launches    600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --ncu
run ncu_tile    400 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_tile -s 2 -c 1 -o gpurun_out/r2g_ncu_tile python tools/bwd_variants.py
run profile     300 python tools/profile_step.py
