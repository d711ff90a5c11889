#!/bin/bash
mkdir -p gpurun_out
for pipe in 0 1; do
COMB_TS_PIPE=$pipe timeout 300 python scripts/conv_floor.py 2>&1 | grep -v Warn | tee gpurun_out/conv_floor_pipe$pipe.txt
done
