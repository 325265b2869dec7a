export PYTHONPATH=.
echo "== regs"; DKS_SHARED_DM=regs timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
echo "== tmem"; timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4
