#!/bin/bash
# ncu full capture (+ source-level sampling) of one kernel: gpu_ncu.sh <tag> <kernel-regex>
OUT=gpurun_out/${1:-ncu}
K=${2:-explain_}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_stdout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/stdout2.log 2>&1
ls -la $OUT
