#!/bin/bash
# full GPU parity suite + smoke + bench (both arms) + ncu launch list of one bench run
bash scripts/gpu_tests.sh
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'], 'profiled ms', round(d['profiled_ms_per_step'],3))
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
print('conv TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3), d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -45 gpurun_out/launches_summary.txt
