export PYTHONPATH=.
for g in 1 0; do echo "== DKS_GRAPH=$g"; for k in "" "--plan-mode per_instance" "--kernel tcgen05"; do
DKS_GRAPH=$g timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $k 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$k', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], round(d['e2e']['value']), d['gpu_launches'])"; done; done
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
