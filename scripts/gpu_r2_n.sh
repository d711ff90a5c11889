#!/bin/bash
# round 2, call N: pipeline ablation of the tcgen05 gather-GEMM
mkdir -p gpurun_out
timeout 600 python scripts/conv_ablate.py > gpurun_out/conv_ablate.txt 2>&1; echo "ablate exit $?"; cat gpurun_out/conv_ablate.txt | tail -8
