#!/bin/bash
# round-end evidence (r2): full GPU parity suite + smoke, bench (both arms), ncu launch list, ncu --set full of the four
# dominant conv shapes, compute-sanitizer racecheck / synccheck / memcheck of the atomics- and barrier-based kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests_r2.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/gpu_tests_r2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r2.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
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
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -30 gpurun_out/launches_summary.txt
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
for cfg in "1 16 16 spconv_tr" "2 32 32 spconv_tr" "3 64 64 spconv_ts" "4 128 128 spconv_ts"; do
set -- $cfg
timeout 300 $NCU -k regex:$4 -s 3 -c 1 -f -o gpurun_out/prof_r2_conv$2 python scripts/conv_one.py $1 $2 $3 > gpurun_out/ncu_conv$2.log 2>&1; echo "ncu conv $2x$3 exit $?"
done
for tool in racecheck synccheck memcheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/racecheck_target.py > gpurun_out/sanitizer_r2_$tool.log 2>&1; echo "$tool exit $?"; tail -3 gpurun_out/sanitizer_r2_$tool.log
done
ls -la gpurun_out/*.ncu-rep
