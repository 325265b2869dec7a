#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the fused kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== smoke" | tee $OUT/smoke.log
timeout 300 python __graft_entry__.py 2>&1 | tail -5 | tee -a $OUT/smoke.log
echo "== bench" | tee $OUT/bench.log
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee -a $OUT/bench.log
if [ -n "$KERNELS" ]; then
  for k in $KERNELS; do
    echo "== bench --kernel $k" | tee -a $OUT/bench.log
    timeout 600 python bench.py --steps 20 --warmup 5 --kernel $k --no-cpu-baseline 2>&1 | tail -2 | tee -a $OUT/bench.log
  done
fi
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches_stdout.log 2>&1
echo "== ncu full capture of the fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_ -s 2 -c 2 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_stdout.log 2>&1
ls -la $OUT
