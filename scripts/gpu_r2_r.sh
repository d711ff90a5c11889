#!/bin/bash
# round 2, call R: per-shape kernel choice (conv_tr for Cin <= 32): spconv + backbone tests, bench auto vs ts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_backbone.py -m gpu -q -x --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_conv_auto.log 2>&1; rc=$?; echo "== spconv+backbone (auto) exit $rc"; tail -3 gpurun_out/test_conv_auto.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_conv_auto.log | head -30; fi
for impl in auto ts; do
if [ $impl = auto ]; then unset COMB_CONV_IMPL; else export COMB_CONV_IMPL=$impl; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err; echo "bench impl=$impl exit $?"; tail -2 gpurun_out/bench_$impl.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$impl.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
