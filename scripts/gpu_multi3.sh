#!/bin/bash
# 8-GPU box, final validation: default collective at N = 8, 4, 2 (gated timed region), l1 + API tests on one GPU of the box.
TAG=${1:-r2m8c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PORT=29700
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
    print("${name} FAILED", e); print(open("$OUT/bench_${name}.log").read()[-1500:])
PY
}
run 8 n8_flags_1 DKS_X=0
run 8 n8_flags_2 DKS_X=0
run 8 n8_nccl DKS_BENCH_NCCL=1
run 4 n4_flags DKS_X=0
run 2 n2_flags DKS_X=0
echo "== l1 + API tests"
timeout 900 python -m pytest tests/test_gpu_l1.py tests/test_gpu_api.py -m gpu -q --timeout 600 2>&1 | tail -3 | tee $OUT/pytest_l1_api.log
echo "== l1 throughput (configs[2] shape, 2048 instances)"
timeout 300 python scripts/gpu_l1_profile.py 2048 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-mode 2>/dev/null | tail -1 > $OUT/bench_n1.json
python -c "
import json; d=json.load(open('$OUT/bench_n1.json')); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value']); [print(k, round(v['value']), v['e2e_reference_default_kwargs']) for k,v in d['other_configs'].items()]"
ls $OUT | wc -l
