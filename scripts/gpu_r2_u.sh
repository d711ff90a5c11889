#!/bin/bash
# round 2, call U: programmatic dependent launch of the conv kernels: parity (incl. graph replay), bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_backbone.py tests/test_gpu_train_fused.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_pdl.log 2>&1; rc=$?; echo "== spconv+backbone+train exit $rc"; tail -3 gpurun_out/test_pdl.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_pdl.log | head -30; fi
for pdl in 1 0 1 0; do
COMB_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl exit $?"; tail -2 gpurun_out/bench_pdl$pdl.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl$pdl.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
PY
done
