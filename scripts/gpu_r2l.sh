#!/bin/bash
# Profile of the 129..1024-group path at the configs[3] singleton shape: launch list (time per kernel), then one
# `ncu --set full` capture of the coalition kernel and the float64 product.  Short: 2.4 GPU-minutes were left.
OUT=gpurun_out/${1:-r2l}; mkdir -p $OUT
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_wide.csv \
    python scripts/gpu_wide_profile.py 2048 > $OUT/launches_run.log 2>&1
tail -2 $OUT/launches_run.log | cut -c1-300
timeout 70 ncu --set full --clock-control none --import-source on -k regex:"wide_beta_kernel|explain_shared_tmem_kernel" -s 3 -c 3 \
    -o $OUT/wide_full python scripts/gpu_wide_profile.py 2048 > $OUT/full_run.log 2>&1
tail -2 $OUT/full_run.log | cut -c1-300; ls -la $OUT
