#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() { timeout 900 $NCU "$@" python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_last.log 2>&1; echo "exit $?"; }
run -k regex:spconv_tc -s 84 -c 1 -o gpurun_out/prof_conv16 -f
run -k regex:spconv_tc -s 95 -c 1 -o gpurun_out/prof_conv64 -f
ls -la gpurun_out/*.ncu-rep
