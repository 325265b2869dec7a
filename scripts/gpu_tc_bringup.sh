#!/bin/bash
# tcgen05 bring-up: accumulator-tile test under both descriptor variants, then the full GPU suite and bench.
OUT=gpurun_out/${1:-tc1}
mkdir -p $OUT
for v in 0 1; do
  echo "== accumulator tile test, DKS_TC_DESC_VARIANT=$v" | tee -a $OUT/bringup.log
  DKS_TC_DESC_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k accumulator_tile 2>&1 | tail -25 | tee -a $OUT/bringup.log
done
echo "== full gpu suite" | tee -a $OUT/bringup.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a $OUT/bringup.log
for k in simt tcgen05; do
  echo "== bench --kernel $k" | tee -a $OUT/bench.log
  timeout 600 python bench.py --steps 20 --warmup 5 --kernel $k --no-cpu-baseline 2>&1 | tail -2 | tee -a $OUT/bench.log
done
