export PYTHONPATH=.
for i in 1 2; do timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], round(d['e2e']['value']), d['gpu_launches'])"; done
DKS_SHARED_DM=regs timeout 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('regs', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
DKS_SHARED_DM=regs timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "shared or adult or config" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
