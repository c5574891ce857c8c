#!/bin/bash
# Profiling recipe behind profiles/r01_*: run from the repo root on a B200 box (gpurun).
#  1. plain run of the same command (must exit 0; its numbers are the only bench numbers)
#  2. launch list (per-launch durations, cold cache, serialised): kernels' SHARE of a step
#  3. ncu --set full of seed_lookup_kernel and of the random-sector microkernel (DRAM bytes per launch)
set -u
O=gpurun_out; mkdir -p $O
CMD="python bench.py --coverage 2.7 --steps 1 --warmup 1 --no-cpu-baseline"     # one 32-Mbase batch per step
MR_BENCH_WATCHDOG=200 timeout 240 $CMD > $O/prof_plain.json 2> $O/prof_plain.err || { echo "plain run failed"; tail $O/prof_plain.err; exit 1; }
MR_BENCH_WATCHDOG=500 timeout 560 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/prof_launches.csv $CMD > $O/prof_ncu1.log 2>&1
MR_BENCH_WATCHDOG=800 timeout 860 ncu --set full --clock-control none --import-source on -k 'regex:seed_lookup_kernel|random_gather_kernel' -c 6 -f -o $O/prof_seed_full $CMD > $O/prof_ncu2.log 2>&1
ncu -i $O/prof_seed_full.ncu-rep --page raw --csv > $O/prof_seed_full_raw.csv 2> /dev/null
ls -la $O/prof_*
