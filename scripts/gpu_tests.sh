#!/bin/bash
# Runs the GPU parity suite file by file (a hang in one file cannot take the others down).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in voxelize rulebook gridindex dense_boxes spconv backbone; do
  if [ -f tests/test_gpu_$f.py ]; then
    timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
    echo "== $f exit $?" | tee -a gpurun_out/summary.txt
    tail -5 gpurun_out/test_$f.log
  fi
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/smoke.log
