#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
for m in 1 0; do
  COMB_DENSE_SCATTER=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_ds$m.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ds$m.json'))
print('scatter=$m value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],4))
PY
done
done
