export PYTHONPATH=.
mkdir -p gpurun_out/ab
for mode in double float; do
  echo "== pmat $mode"
  DKS_PMAT=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ab/l_$mode.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  grep -E "wls_pmat|explain_shared" gpurun_out/ab/l_$mode.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -4
  DKS_PMAT=$mode timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
