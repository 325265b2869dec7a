#!/bin/bash
# Plans of more than 128 groups (sixteen-word rows + projection solve): the new GPU tests first (stop early if they
# fail), then the default bench line (its other_configs now carry the configs[3] singleton reading), then the full suite.
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_wide.py "tests/test_gpu_baseline_shapes.py::test_config3_singleton_reading_1024_groups_shared_plan" -x -q -s --timeout 200 2>&1 | tail -60 > $OUT/new_tests.log
RC=${PIPESTATUS[0]}
tail -25 $OUT/new_tests.log
if [ $RC -ne 0 ]; then echo "NEW TESTS FAILED rc=$RC"; exit 1; fi
timeout 240 python bench.py --steps 20 --warmup 5 2>$OUT/bench_default.err | tail -1 > $OUT/bench_default.json
python -c "
import json; d=json.load(open('$OUT/bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'per_instance', d['per_instance']['value']); [print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'), v.get('error')) for k,v in d['other_configs'].items()]" || tail -5 $OUT/bench_default.err
timeout 330 python -m pytest tests -m gpu -q -x --timeout 300 --deselect tests/test_gpu_wide.py 2>&1 | tail -4 | tee $OUT/pytest.log
