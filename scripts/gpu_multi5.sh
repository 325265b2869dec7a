#!/bin/bash
# 8-GPU box: in-kernel push vs separate push kernel (both followed by the engine's flag exchange), aligned + gated timed region.
OUT=gpurun_out/${1:-r2m8d}; mkdir -p $OUT
PORT=29900
run() {
  N=$1; name=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${name}.log 2>&1
  PORT=$((PORT+1))
  grep '^{"metric"' $OUT/bench_${name}.log | tail -1 > $OUT/bench_${name}.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${name}.json").read())
    print("${name} N=$N", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "steps", d.get("step_ms_rank0"))
except Exception as e:
    print("${name} FAILED", e); print(open("$OUT/bench_${name}.log").read()[-1500:])
PY
}
run 8 n8_inkernel DKS_X=0
run 8 n8_pushkernel_1 DKS_PUSH_IN_KERNEL=0
run 8 n8_pushkernel_2 DKS_PUSH_IN_KERNEL=0
run 8 n8_nccl DKS_BENCH_NCCL=1
run 4 n4_pushkernel DKS_PUSH_IN_KERNEL=0
run 4 n4_inkernel DKS_X=0
ls $OUT | wc -l
