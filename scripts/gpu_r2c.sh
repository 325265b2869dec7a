#!/bin/bash
# Round-2 GPU session C: fused kernel knobs (batch, instances per pass), full default bench line, launch list, tests.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run_bench() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode --no-other-configs 2>&1 | tail -1 > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read())
    print("$name", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "sustained", d.get("sustained", {}).get("ms_per_step"))
except Exception as e:
    print("$name FAILED", e, open("$OUT/bench_$name.json").read()[-1500:])
PY
}
run_bench b16 DKS_X=0
run_bench b8 DKS_FUSED_B=8
run_bench b32 DKS_FUSED_B=32
run_bench ni2_b16 DKS_FUSED_NI=2
run_bench ni2_b8 DKS_FUSED_NI=2 DKS_FUSED_B=8
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 | tee $OUT/pytest.log
echo "== NI=2 parity"
DKS_FUSED_NI=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -q -k "adult or shared or golden or randomized" --timeout 600 2>&1 | tail -5 | tee $OUT/pytest_ni2.log
echo "== full default bench line"
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_default.json; python -c "
import json; d=json.load(open('$OUT/bench_default.json')); print({k: d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d.get('per_instance',{}).get('value'), d.get('sustained',{}).get('ms_per_step')); print(json.dumps(d.get('other_configs'), indent=1)[:2500])"
echo "== reference arm"
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 2) 2>&1 | tail -5 | cut -c1-400 | tee $OUT/bench_reference.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-other-configs > $OUT/ncu_launches_stdout.log 2>&1
echo "== cfg4 on one GPU (1.25 M instances: one rank's share)"
timeout 600 python scripts/gpu_cfg4_multi.py 1250000 2>&1 | tail -1 | tee $OUT/cfg4_1gpu.json
ls -la $OUT
