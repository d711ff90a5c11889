#!/bin/bash
# round 2, call D: pipelined gather with 8-byte direct-to-fragment loads (parity, then A/B bench), BN kernels vs torch
mkdir -p gpurun_out
COMB_TS_PIPE=1 timeout 300 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv_pipe1.log 2>&1; rc=$?; echo "== spconv (pipelined gather) exit $rc"; tail -5 gpurun_out/test_spconv_pipe1.log
if [ $rc = 0 ]; then
for pipe in 0 1; do
COMB_TS_PIPE=$pipe timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_pipe$pipe.json 2> gpurun_out/bench_pipe$pipe.err; echo "bench pipe=$pipe exit $?"; tail -2 gpurun_out/bench_pipe$pipe.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pipe$pipe.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
fi
timeout 600 python -m pytest tests/test_gpu_bn.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_bn.log 2>&1; echo "== bn exit $?"; tail -15 gpurun_out/test_bn.log
