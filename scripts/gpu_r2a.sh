#!/bin/bash
# Round-2 GPU session A: pipe probe, parity tests (incl. full BASELINE shapes), bench variants of the fused kernel, ncu.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== quad probe"; timeout 120 scripts/probes/quad_probe > $OUT/quad_probe.txt 2>&1; tail -30 $OUT/quad_probe.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 2>&1 | tail -80 | tee $OUT/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -5 | tee $OUT/smoke.log
run_bench() {  # name, env...
  name=$1; shift
  echo "== bench $name"
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read())
    print("$name", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
except Exception as e:
    print("$name FAILED", e, open("$OUT/bench_$name.json").read()[-2000:])
PY
}
run_bench fused_ni1_w20 DKS_X=0
run_bench unfused DKS_FUSED=0
run_bench fused_ni2_w20 DKS_FUSED_NI=2
run_bench fused_ni1_w16 DKS_FUSED_WARPS=16
run_bench fused_ni2_w16 DKS_FUSED_NI=2 DKS_FUSED_WARPS=16
run_bench fused_ni1_w20_b16 DKS_FUSED_B=16
run_bench fused_ni2_w16_b16 DKS_FUSED_NI=2 DKS_FUSED_WARPS=16 DKS_FUSED_B=16
echo "== bench per_instance"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --plan-mode per_instance 2>&1 | tail -1 > $OUT/bench_per_instance.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches_stdout.log 2>&1
echo "== ncu full capture of the fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_shared_fused -s 2 -c 1 -f -o $OUT/prof_fused \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_stdout.log 2>&1
echo "== compute-sanitizer memcheck (smoke)"
timeout 600 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck.log python __graft_entry__.py > $OUT/sanitizer_memcheck_stdout.log 2>&1
tail -5 $OUT/sanitizer_memcheck.log
ls -la $OUT
