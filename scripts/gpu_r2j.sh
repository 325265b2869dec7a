#!/bin/bash
# Round-2 final confirmation on the shipped build: GPU tests, smoke, the two bench arms the driver runs.
OUT=gpurun_out/${1:-r2j}; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest.log
timeout 120 python __graft_entry__.py 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $OUT/bench_default.json
python -c "
import json; d=json.load(open('$OUT/bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'per_instance', d['per_instance']['value'], 'sustained', d['sustained']['ms_per_step'], d['step_ms_rank0']); [print(k, round(v['value']), v['e2e_reference_default_kwargs']) for k,v in d['other_configs'].items()]"
timeout 150 python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 > $OUT/bench_reference.json; cut -c1-260 $OUT/bench_reference.json
