#!/bin/bash
# round 2, call W: conv_ts single-tile passes split their K range over both halves of the pipeline (COMB_TS_SPLIT);
# velocity-head decode; bench A/B on one box
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
run() { echo "--- split=$1 case $2 $3 $4 $5"; COMB_TS_SPLIT=$1 timeout -s KILL 25 python scripts/ts_split_diag.py $2 $3 $4 $5 2>&1 | tail -1 | grep -q "done err [0-9.]*e-0[5-9]" && echo ok || { echo FAILED; return 1; }; }
run 1 128 128 3 3000 && run 1 64 64 27 3000 && run 3 64 64 27 3000 && run 1 128 128 27 13000 || exit 1
timeout -s KILL 400 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_center_decode.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/test_split_a.log 2>&1; rc=$?; echo "== spconv + decode tests exit $rc"; tail -3 gpurun_out/test_split_a.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_split_a.log | head -30; exit 1; fi
for sp in 0 1 0 1; do
COMB_TS_SPLIT=$sp timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_split$sp.json 2> gpurun_out/bench_split$sp.err; echo "bench split=$sp exit $?"; tail -2 gpurun_out/bench_split$sp.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/bench_split$sp.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
