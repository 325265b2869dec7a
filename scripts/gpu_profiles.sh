#!/bin/bash
# round-end evidence: bench lines, ncu launch lists, ncu --set full captures of the dominant kernels
OUT=gpurun_out/${1:-final}
export PYTHONPATH=.
mkdir -p $OUT
timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > $OUT/bench_auto.json
timeout 600 python bench.py --steps 30 --warmup 5 --kernel tcgen05 --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_tcgen05.json
timeout 600 python bench.py --steps 30 --warmup 5 --plan-mode per_instance --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_per_instance.json
timeout 600 python scripts/gpu_cfg2.py 32768 shared cfg2 2>&1 | tail -1 > $OUT/cfg2_shared.json
timeout 600 python scripts/gpu_cfg2.py 2048 per_instance cfg2 2>&1 | tail -1 > $OUT/cfg2_per_instance.json
timeout 600 python scripts/gpu_cfg2.py 16384 shared cfg3 2>&1 | tail -1 > $OUT/cfg3_grouped_shared.json
timeout 600 python scripts/gpu_cfg2.py 16384 shared cfg4 2>&1 | tail -1 > $OUT/cfg4_shared.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_auto.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_per_instance.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --plan-mode per_instance > /dev/null 2>&1
for spec in "explain_shared auto shared" "wls_pmat auto shared" "explain_tcgen05 tcgen05 shared" "sample_plans auto per_instance"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o $OUT/prof_$1 \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --kernel $2 --plan-mode $3 > $OUT/ncu_$1.log 2>&1
  # the reports are ~18 MB each (gpurun brings back at most 64 MiB): keep the raw and source pages as CSV
  ncu -i $OUT/prof_$1.ncu-rep --page raw --csv > $OUT/raw_$1.csv 2>/dev/null
  ncu -i $OUT/prof_$1.ncu-rep --page source --csv > $OUT/src_$1.csv 2>/dev/null
  rm -f $OUT/prof_$1.ncu-rep
done
ls -la $OUT
