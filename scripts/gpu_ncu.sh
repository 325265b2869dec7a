#!/bin/bash
# ncu full capture (+ source-level sampling) of the fused kernel only
OUT=gpurun_out/${1:-ncu}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_ -s 3 -c 1 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_stdout.log 2>&1
ls -la $OUT
