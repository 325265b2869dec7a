#!/bin/bash
# quick iteration: parity tests + bench of the default kernel (no CPU baseline)
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
for k in auto tcgen05; do timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --kernel $k 2>&1 | tail -1 | tee -a $OUT/bench.log; done
