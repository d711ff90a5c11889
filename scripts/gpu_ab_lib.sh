#!/bin/bash
# A/B of two builds of the library in one box: libcomb200.so (base) vs libcomb200_var.so (variant)
mkdir -p gpurun_out
L=com_b200/lib
cp $L/libcomb200.so $L/base.keep
for i in 1 2; do
for m in base var; do
  if [ $m = base ]; then cp $L/base.keep $L/libcomb200.so; else cp $L/libcomb200_var.so $L/libcomb200.so; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$m.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$m.json'))
print('$m value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1), 'conv', round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],4), {k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
done
cp $L/libcomb200_var.so $L/libcomb200.so
timeout 600 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x -k "bf16" -p no:cacheprovider 2>&1 | tail -2
cp $L/base.keep $L/libcomb200.so
