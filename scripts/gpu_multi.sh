#!/bin/bash
# multi-GPU session: bench at N GPUs with the three collectives, the SPMD e2e path, optional cfg4 at full size.
# usage: bash scripts/gpu_multi.sh N tag [cfg4_instances]
N=${1:-2}; TAG=${2:-r2m$N}; CFG4=${3:-0}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PORT=29533
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_$name.log 2>&1
  PORT=$((PORT+1))
  grep '^{"metric"' $OUT/bench_$name.log | tail -1 > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read())
    print("$name N=$N", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "|", d["config"]["collective"][:70])
except Exception as e:
    print("$name FAILED", e); print(open("$OUT/bench_$name.log").read()[-1500:])
PY
}
run push_flags DKS_X=0
run push_symm DKS_BENCH_SYMM_BARRIER=1
run nccl DKS_BENCH_NCCL=1
run push_flags_again DKS_X=0
if [ "$CFG4" != "0" ]; then
  echo "== cfg4 at $CFG4 instances on $N GPUs"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      scripts/gpu_cfg4_multi.py $CFG4 2>&1 | grep '^{' | tail -1 | tee $OUT/cfg4_${N}gpu.json
fi
if [ "$N" = "2" ]; then
  echo "== cpu scaling probe (reference arm diagnostics)"
  timeout 600 python scripts/cpu_scaling_probe.py 2>&1 | tail -6 | tee $OUT/cpu_scaling.log
fi
ls -la $OUT
