#!/bin/bash
# A/B of the two tcgen05 sparse-conv kernels: parity tests on the default (ts), then the bench with each.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "bf16" > gpurun_out/test_spconv_ts.log 2>&1; echo "== spconv(ts) exit $?"; tail -15 gpurun_out/test_spconv_ts.log
for impl in ts ss; do
  COMB_CONV_IMPL=$impl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err; echo "bench $impl exit $?"; tail -3 gpurun_out/bench_$impl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$impl.json'))
    print('$impl value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
    print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
    print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
    print('conv TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3))
except Exception as e:
    print('no bench json', e)
PY
done
