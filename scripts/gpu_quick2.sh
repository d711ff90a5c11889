#!/bin/bash
# tests named on the command line + bench + ncu launch list
mkdir -p gpurun_out
for f in "$@"; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$f.log 2>&1; echo "== $f exit $?"; tail -3 gpurun_out/test_$f.log
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; grep "index_\|vox_\|nbrmap\|dense\|permute" gpurun_out/launches_summary.txt
