#!/bin/bash
# round 2, call C: gather-loop variants (0 = r1, 1 = pipelined + register budget, 2 = r1 loop + the same budget),
# ncu --set full of the pipelined 64x64 kernel, compute-sanitizer racecheck/synccheck of the atomics-based kernels
mkdir -p gpurun_out
for pipe in 0 2 1; do
COMB_TS_PIPE=$pipe timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_pipe$pipe.json 2> gpurun_out/bench_pipe$pipe.err; echo "bench pipe=$pipe exit $?"; tail -2 gpurun_out/bench_pipe$pipe.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pipe$pipe.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'conv',round(d['breakdown_ms_per_step']['spconv_fwd_bf16'],3))
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
COMB_TS_PIPE=1 timeout 300 $NCU -k regex:"spconv_ts_kernel.*\)64, .*\)64" -s 12 -c 1 -o gpurun_out/prof_ts64_pipe1 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ts64_pipe1.log 2>&1; echo "ncu pipe1 exit $?"
COMB_TS_PIPE=0 timeout 300 $NCU -k regex:"spconv_ts_kernel.*\)64, .*\)64" -s 12 -c 1 -o gpurun_out/prof_ts64_pipe0 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ts64_pipe0.log 2>&1; echo "ncu pipe0 exit $?"
for tool in racecheck synccheck; do
timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/racecheck_target.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool exit $?"; tail -4 gpurun_out/sanitizer_$tool.log
done
