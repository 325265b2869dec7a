#!/bin/bash
# per-instance plan mode: bench line + launch list (kernel shares)
mkdir -p gpurun_out
timeout 300 python bench.py --plan-mode per_instance --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_perinst.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/perinst_launches.csv \
    python bench.py --plan-mode per_instance --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, json, collections
line = json.load(open("gpurun_out/bench_perinst.json"))
print({k: line[k] for k in ("value", "ms_per_step")}, line["e2e"]["value"], line["roofline"]["kernel_ms"])
rows = [r for r in csv.reader(l for l in open("gpurun_out/perinst_launches.csv") if not l.startswith("=="))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) > vi:
        try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
        except ValueError: pass
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
