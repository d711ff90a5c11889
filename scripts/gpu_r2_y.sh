#!/bin/bash
# round 2, call Y: conv_ts gather with one 32-byte load per row (COMB_TS_WIDE): parity, then bench A/B on one box
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
run() { echo "--- $1 case $2 $3 $4 $5"; env $1 timeout -s KILL 25 python scripts/ts_split_diag.py $2 $3 $4 $5 2>&1 | tail -1 | grep "done err [0-9.]*e-0[5-9]" || { echo FAILED; return 1; }; }
run COMB_TS_WIDE=1 64 64 27 3000 && run COMB_TS_WIDE=1 128 128 27 13000 && run "COMB_TS_WIDE=1 COMB_CONV_IMPL=ts" 32 32 27 3000 && run "COMB_TS_WIDE=1 COMB_CONV_IMPL=ts" 16 16 27 3000 && run COMB_TS_WIDE=1 128 128 3 3000 || exit 1
COMB_TS_WIDE=1 timeout -s KILL 400 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_backbone.py tests/test_gpu_train_fused.py tests/test_gpu_reference_dropin.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/test_wide.log 2>&1; rc=$?; echo "== conv-dependent tests under COMB_TS_WIDE=1 exit $rc"; tail -3 gpurun_out/test_wide.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_wide.log | head -30; exit 1; fi
for cfg in "0 auto" "1 auto" "0 auto" "1 auto" "1 ts"; do
set -- $cfg
if [ $2 = ts ]; then export COMB_CONV_IMPL=ts; else unset COMB_CONV_IMPL; fi
COMB_TS_WIDE=$1 timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_wide$1_$2.json 2> gpurun_out/bench_wide$1_$2.err; echo "bench wide=$1 impl=$2 exit $?"; tail -2 gpurun_out/bench_wide$1_$2.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/bench_wide$1_$2.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
