#!/bin/bash
# quick iteration: parity tests + bench of the default kernel (no CPU baseline)
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.log
