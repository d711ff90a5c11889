#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -W ignore scripts/full_step.py > gpurun_out/full_step.json 2> gpurun_out/full_step.err; echo "full step exit $?"; grep -v "^frame\|CUDAEvent\|Warning" gpurun_out/full_step.err | tail -15; cat gpurun_out/full_step.json | cut -c1-3000
