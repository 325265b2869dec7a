#!/bin/bash
OUT=gpurun_out/${1:-variants}
mkdir -p $OUT
for lib in build/variants/libdks_*.so; do
  name=$(basename $lib .so)
  echo -n "$name " | tee -a $OUT/variants.log
  DKS_LIB=$PWD/$lib timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -o "smoke ok.*max rel err vs oracle [0-9.e+-]*\|Error.*\|rel err.*" | head -1 | tr '\n' ' ' | tee -a $OUT/variants.log
  DKS_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('kernel_ms',round(d['roofline']['kernel_ms'],4),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
except Exception as e: print('FAILED',e)" | tee -a $OUT/variants.log
done
