#!/bin/bash
OUT=gpurun_out/${1:-ablate}
mkdir -p $OUT
for a in ${ABL:-0 31 63 95 159 223 479 511}; do
  echo -n "ablate=$a " | tee -a $OUT/ablate.log
  DKS_TC_ABLATE=$a timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('kernel_ms',round(d['roofline']['kernel_ms'],4),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))" | tee -a $OUT/ablate.log
done
