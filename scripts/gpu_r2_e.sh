#!/bin/bash
# round 2, call E: tile assignment (blocked / strided) x ring depths (more or less L1 left to the gather); fused train step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv_blocked.log 2>&1; rc=$?; echo "== spconv (blocked default) exit $rc"; tail -3 gpurun_out/test_spconv_blocked.log
for cfg in "0 8 8" "1 8 8" "0 4 4" "1 4 4" "1 4 2"; do
set -- $cfg
COMB_TS_BLOCKED=$1 COMB_TS_NI=$2 COMB_TS_NB=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; echo "bench blocked=$1 ni=$2 nb=$3 exit $?"; tail -2 gpurun_out/bench_cfg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
timeout 600 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/test_train_fused.log 2>&1; echo "== train_fused exit $?"; tail -30 gpurun_out/test_train_fused.log
