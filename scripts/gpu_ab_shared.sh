export PYTHONPATH=.
for i in 1 2; do timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
