#!/bin/bash
# round-end evidence (r2, after the split single-tile passes of conv_ts and the velocity-head decode): full GPU parity
# suite + smoke, bench (both arms), ncu launch list, synccheck / memcheck / racecheck of the sanitizer target
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/gpu_tests_r2.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/gpu_tests_r2.log; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests_r2.log | head
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r2.log
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'], 'profiled ms', round(d['profiled_ms_per_step'],3))
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
print('conv TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3), d.get('cpu_baseline'))
PY
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -12 gpurun_out/launches_summary.txt
for tool in synccheck memcheck racecheck; do
timeout -s KILL 240 compute-sanitizer --tool $tool --print-limit 20 python scripts/racecheck_target.py > gpurun_out/sanitizer_r2_$tool.log 2>&1; echo "$tool exit $?"; tail -3 gpurun_out/sanitizer_r2_$tool.log
done
