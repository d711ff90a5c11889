#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bn.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_bn.log 2>&1; echo "== bn exit $?"; tail -2 gpurun_out/test_bn.log
timeout 600 python -m pytest tests/test_gpu_train_fused.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/test_train_fused.log 2>&1; echo "== train_fused (graph) exit $?"; grep -v "^$" gpurun_out/test_train_fused.log | grep -E "smooth variant|fused train step|Error|passed|failed|^E " | cut -c1-400 | tail -12
timeout 300 python scripts/train_bench.py --prof > gpurun_out/train_prof.log 2>&1; echo "train prof exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/train_prof.log | grep -v "^-" | cut -c1-75,150-260 | tail -32
