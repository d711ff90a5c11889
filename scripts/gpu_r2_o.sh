#!/bin/bash
# round 2, call O: row-per-thread tcgen05 kernel (conv_tr.cu): parity, then per-layer times against conv_ts
mkdir -p gpurun_out
COMB_CONV_IMPL=tr timeout 400 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv_tr.log 2>&1; rc=$?; echo "== spconv (tr) exit $rc"; tail -3 gpurun_out/test_spconv_tr.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_spconv_tr.log | head -30; exit 1; fi
for impl in tr ts; do
COMB_CONV_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err; echo "bench impl=$impl exit $?"; tail -2 gpurun_out/bench_$impl.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$impl.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
