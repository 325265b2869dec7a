#!/bin/bash
# 2-GPU check of the aligned, gated timed region
OUT=gpurun_out/${1:-r2m2b}; mkdir -p $OUT
PORT=29800
for rep in 1 2 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_$rep.log 2>&1
  PORT=$((PORT+1))
  grep '^{"metric"' $OUT/bench_n2_$rep.log | tail -1 > $OUT/bench_n2_$rep.json
  python -c "
import json
try:
    d=json.load(open('$OUT/bench_n2_$rep.json')); print('N=2 rep $rep', d['ms_per_step'], d['value'], d['e2e']['value'], d['step_ms_rank0'])
except Exception as e: print('FAILED', e); print(open('$OUT/bench_n2_$rep.log').read()[-1500:])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
