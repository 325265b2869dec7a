export PYTHONPATH=.
for k in "--kernel tcgen05" "--plan-mode per_instance" ""; do
timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $k 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$k', d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"; done
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
