#!/bin/bash
# round 2, call M: larger A stages (SC chunks of each tile per stage) and sleep-free background waits
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent or alternate" > gpurun_out/test_spconv_sc.log 2>&1; rc=$?; echo "== spconv (SC stages) exit $rc"; tail -3 gpurun_out/test_spconv_sc.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout" gpurun_out/test_spconv_sc.log | head -20; exit 1; fi
for cfg in "2 0" "0 0" "0 2" "2 2"; do
set -- $cfg
COMB_TS_SC=$1 COMB_TS_PIPE=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; echo "bench sc=$1 (0 = default 3/2) pipe=$2 exit $?"; tail -2 gpurun_out/bench_cfg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
