#!/bin/bash
# Round-2 GPU session B: fused kernel v2 (compile-time N, trimmed per-row work): tests, bench variants, ncu, sanitizer.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (selected)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 | tee $OUT/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -3 | tee $OUT/smoke.log
run_bench() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode --no-other-configs 2>&1 | tail -1 > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$name.json").read())
    print("$name", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "launches", d["gpu_launches"], "sustained", d.get("sustained", {}).get("ms_per_step"))
except Exception as e:
    print("$name FAILED", e, open("$OUT/bench_$name.json").read()[-1500:])
PY
}
run_bench fused DKS_X=0
run_bench fused_b16 DKS_FUSED_B=16
run_bench fused_w16 DKS_FUSED_WARPS=16
run_bench unfused DKS_FUSED=0
echo "== full default bench line (both plan modes, cpu baseline)"
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_default.json; cut -c1-600 $OUT/bench_default.json
echo "== ncu full capture of the fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_shared_fused -s 2 -c 1 -f -o $OUT/prof_fused \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-mode --no-other-configs > $OUT/ncu_full_stdout.log 2>&1
echo "== compute-sanitizer memcheck (smoke)"
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck.log python __graft_entry__.py > $OUT/sanitizer_memcheck_stdout.log 2>&1
tail -4 $OUT/sanitizer_memcheck.log; tail -2 $OUT/sanitizer_memcheck_stdout.log
ls -la $OUT
