#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bn.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_bn.log 2>&1; echo "== bn exit $?"; tail -4 gpurun_out/test_bn.log
timeout 600 python -m pytest tests/test_gpu_train_fused.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/test_train_fused.log 2>&1; echo "== train_fused exit $?"; grep -v "^$" gpurun_out/test_train_fused.log | grep -E "fused train step|Error|passed|failed|^E " | cut -c1-900 | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
t=d['train_sparse_part']
for k,v in t.items():
    if isinstance(v,dict) and 'ms_per_step' in v: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='layers'})
print({k:(round(v['ms_per_launch']*1e3,1), round(v['tflops'],1)) for k,v in t['wgrad_tcgen05']['layers'].items()})
PY
