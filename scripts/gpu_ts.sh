#!/bin/bash
# parity tests of the bf16 conv, pipeline trace, bench (ts only)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "bf16" > gpurun_out/test_spconv_ts.log 2>&1; rc=$?; echo "== spconv(ts) exit $rc"; tail -5 gpurun_out/test_spconv_ts.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python scripts/conv_trace_ts.py > gpurun_out/trace_ts.log 2>&1; echo "trace exit $?"; grep "==\|mma\|gather\|epilogue" gpurun_out/trace_ts.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_ts.json 2> gpurun_out/bench_ts.err; echo "bench exit $?"; tail -3 gpurun_out/bench_ts.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ts.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
print('conv TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3))
PY
