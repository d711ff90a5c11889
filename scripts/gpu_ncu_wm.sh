#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:"spconv_wm_kernel.*16.*16" -s 12 -c 1 -o gpurun_out/prof_wm16 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_last.log 2>&1; echo "exit $?"
timeout 600 $NCU -k regex:"spconv_wm_kernel.*32.*32" -s 10 -c 1 -o gpurun_out/prof_wm32 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_last2.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
