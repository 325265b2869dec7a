#!/bin/bash
# Round-2 final GPU session: tests, smoke, the bench lines the driver will run, ncu captures of every hot kernel, sanitizers.
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee $OUT/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $OUT/bench_reference.json; cut -c1-300 $OUT/bench_reference.json
echo "== bench default"
(time timeout 900 python bench.py --steps 20 --warmup 5) > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -3 $OUT/bench_default.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "per_instance", d.get("per_instance", {}).get("value"), "sustained", d.get("sustained", {}).get("ms_per_step"))
print(json.dumps(d["roofline"])[:900])
PY
echo "== bench per_instance as primary"
timeout 600 python bench.py --steps 20 --warmup 5 --plan-mode per_instance --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 > $OUT/bench_per_instance.json; cut -c1-200 $OUT/bench_per_instance.json
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_shared.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-other-configs --no-other-mode > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $OUT/launches_per_instance.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-other-configs --no-other-mode --plan-mode per_instance > /dev/null 2>&1
echo "== ncu full: fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_shared_fused -s 2 -c 1 -f -o $OUT/prof_fused \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-mode --no-other-configs > /dev/null 2>&1
echo "== ncu full: l1 kernels"
timeout 900 ncu --set full --clock-control none -k regex:l1_ -s 2 -c 2 -f -o $OUT/prof_l1 python scripts/gpu_l1_profile.py 1024 > $OUT/l1_profile_stdout.log 2>&1
echo "== ncu full: per-instance kernels (sampler, tcgen05)"
timeout 900 ncu --set full --clock-control none -k regex:"sample_plans|explain_tcgen05" -s 4 -c 2 -f -o $OUT/prof_per_instance \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-mode --no-other-configs --plan-mode per_instance > /dev/null 2>&1
echo "== compute-sanitizer memcheck + racecheck (smoke)"
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck.log python __graft_entry__.py > /dev/null 2>&1; tail -2 $OUT/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --log-file $OUT/sanitizer_racecheck.log python __graft_entry__.py > /dev/null 2>&1; tail -3 $OUT/sanitizer_racecheck.log
ls -la $OUT
