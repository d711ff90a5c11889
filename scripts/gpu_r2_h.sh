#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/train_bench.py --prof > gpurun_out/train_prof.log 2>&1; echo "train prof exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/train_prof.log | grep -v "^-" | cut -c1-75,150-260 | tail -45
