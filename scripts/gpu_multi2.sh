#!/bin/bash
# 8-GPU box: repeated bench runs per collective at N=8 and N=4 (choose the default from a table, not one sample).
TAG=${1:-r2m8b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PORT=29600
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
    print("${name} N=$N", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "steps", d.get("step_ms_rank0"), "|", d["config"]["collective"][:40])
except Exception as e:
    print("${name} FAILED", e); print(open("$OUT/bench_${name}.log").read()[-1200:])
PY
}
for rep in 1 2; do
  run 8 n8_flags_$rep DKS_X=0
  run 8 n8_symm_$rep DKS_BENCH_SYMM_BARRIER=1
  run 8 n8_nccl_$rep DKS_BENCH_NCCL=1
done
run 4 n4_flags DKS_X=0
run 4 n4_symm DKS_BENCH_SYMM_BARRIER=1
run 4 n4_nccl DKS_BENCH_NCCL=1
run 2 n2_flags DKS_X=0
echo "== pool of GPUs in one process + API tests on a multi-GPU box"
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q --timeout 500 2>&1 | tail -3 | tee $OUT/pytest_api_multi_gpu.log
ls $OUT | head -40
