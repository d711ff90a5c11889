#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_fused.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/test_train_fused.log 2>&1; echo "== train_fused (graph) exit $?"; grep -v "^$" gpurun_out/test_train_fused.log | grep -E "smooth variant|fused train step|Error|passed|failed|^E " | cut -c1-600 | tail -20
COMB_TRAIN_GRAPH=0 timeout 600 python -m pytest tests/test_gpu_train_fused.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/test_train_fused_eager.log 2>&1; echo "== train_fused (eager) exit $?"; grep -v "^$" gpurun_out/test_train_fused_eager.log | grep -E "smooth variant|fused train step|Error|passed|failed|^E " | cut -c1-600 | tail -20
timeout 300 python scripts/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train bench (graph) exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/train_bench.log | tail -4
COMB_TRAIN_GRAPH=0 timeout 300 python scripts/train_bench.py > gpurun_out/train_bench_eager.log 2>&1; echo "train bench (eager) exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/train_bench_eager.log | tail -2
timeout 300 python scripts/train_bench.py --module > gpurun_out/train_bench_module.log 2>&1; echo "train bench (module) exit $?"; grep -v "^frame\|CUDAEvent" gpurun_out/train_bench_module.log | tail -2
