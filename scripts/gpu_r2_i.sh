#!/bin/bash
mkdir -p gpurun_out
for f in backbone dense_boxes; do
timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$f.log 2>&1; echo "== $f exit $?"; tail -3 gpurun_out/test_$f.log
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/bench.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'], 'd2h', d['e2e']['d2h_bytes_per_step'], d['e2e'].get('numa'))
t=d['train_sparse_part']
if 'error' in t: print(t)
else:
  for k,v in t.items():
    if isinstance(v,dict) and 'ms_per_step' in v: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='layers'})
  print({k:(round(v['ms_per_launch']*1e3,1), round(v['tflops'],1)) for k,v in t['wgrad_tcgen05']['layers'].items()})
print(json.dumps(d['box_ops'].get('e2e', d['box_ops']))[:900])
print(json.dumps(d['comaug_part'].get('e2e', d['comaug_part']))[:900])
PY
