#!/usr/bin/env bash
# round 2, last single-GPU validation of HEAD: smoke, full suite, configs[1] bench line (with cpu_baseline), teacher-student
# line, launch list, kernel profile.  Outputs under gpurun_out/r2h_*.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2h_$name.txt" 2>&1; echo "rc=$? ($name)"; (grep '^{' "gpurun_out/r2h_$name.txt" || tail -n 3 "gpurun_out/r2h_$name.txt") | cut -c1-260; }
run smoke       300 python __graft_entry__.py --smoke
run suite       900 python -m pytest tests -m gpu -q
run bench_sup   600 python bench.py --steps 20 --warmup 5
run bench_ssod  900 python bench.py --workload ssod --steps 8 --warmup 8 --no-cpu-baseline
run launches    600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --ncu
run profile     300 python tools/profile_step.py
