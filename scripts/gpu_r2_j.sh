#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_center_decode.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_center_decode.log 2>&1; echo "== center_decode exit $?"; grep -E "^E |passed|failed|Error" gpurun_out/test_center_decode.log | cut -c1-300 | tail -25
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_nocpu.json 2> gpurun_out/bench_nocpu.err; echo "bench exit $?"; tail -2 gpurun_out/bench_nocpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nocpu.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],3),'d2h', d['e2e']['d2h_bytes_per_step'])
PY
