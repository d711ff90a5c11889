#!/bin/bash
# ncu --set full of the dominant kernel (one launch per channel width), after the graph warm-up of bench.py
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
for w in 16 64 128; do
  timeout 600 $NCU -k regex:"spconv_ts_kernel.*\)$w, .*\)$w>" -s 12 -c 1 -o gpurun_out/prof_ts$w -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ts$w.log 2>&1; echo "ts$w exit $?"
done
ls -la gpurun_out/*.ncu-rep
