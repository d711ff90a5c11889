#!/bin/bash
for S in 4 2 1; do
  echo "=== COMB_TC_S=$S"
  COMB_TC_S=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_S$S.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_S$S.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'conv ms',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
