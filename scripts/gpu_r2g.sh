#!/bin/bash
# Round-2 session G: staged prep kernel + out-of-line peer push: tests, bench, launch list, l1 ncu capture.
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee $OUT/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -1 | tee $OUT/smoke.log
echo "== bench default"
timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $OUT/bench_default.json
python - <<PY
import json
d = json.loads(open("$OUT/bench_default.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "per_instance", d.get("per_instance", {}).get("value"), "sustained", d.get("sustained", {}).get("ms_per_step"), d["step_ms_rank0"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
PY
echo "== bench again (no extras)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode --no-other-configs 2>/dev/null | tail -1 > $OUT/bench_plain.json
python -c "
import json; d=json.load(open('$OUT/bench_plain.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_shared.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-other-configs --no-other-mode > /dev/null 2>&1
grep -c prep_kernel $OUT/launches_shared.csv
echo "== ncu full: l1 kernels"
timeout 900 ncu --set full --clock-control none -k regex:l1_ -s 2 -c 2 -f -o $OUT/prof_l1 python scripts/gpu_l1_profile.py 1024 > $OUT/l1_profile_stdout.log 2>&1; tail -2 $OUT/l1_profile_stdout.log
echo "== ncu full: prep kernel"
timeout 600 ncu --set full --clock-control none -k regex:prep_kernel -s 3 -c 1 -f -o $OUT/prof_prep \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-mode --no-other-configs > /dev/null 2>&1
ls -la $OUT
