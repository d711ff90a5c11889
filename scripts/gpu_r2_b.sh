#!/bin/bash
# round 2, call B: reference-through-dropins tests (byte-code travels as *.bc), conv parity with the pipelined gather,
# A/B bench of the gather loop (COMB_TS_PIPE=0 / 1).  Every step has a short timeout: a hung kernel must not eat the budget.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv.log 2>&1; rc=$?; echo "== spconv (pipelined gather) exit $rc"; tail -5 gpurun_out/test_spconv.log
PIPE_OK=$rc
COMB_TS_PIPE=0 timeout 900 python -m pytest tests/test_gpu_reference_dropin.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_reference_dropin.log 2>&1; echo "== reference_dropin exit $?"; tail -40 gpurun_out/test_reference_dropin.log
for pipe in 0 1; do
if [ $pipe = 1 ] && [ $PIPE_OK != 0 ]; then echo "skipping pipe=1 bench"; continue; fi
COMB_TS_PIPE=$pipe timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_pipe$pipe.json 2> gpurun_out/bench_pipe$pipe.err; echo "bench pipe=$pipe exit $?"; tail -3 gpurun_out/bench_pipe$pipe.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pipe$pipe.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
