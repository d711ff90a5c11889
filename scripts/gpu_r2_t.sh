#!/bin/bash
# round 2, call T: f3 (COMAug placement) + f4 (bf16 NHWC BEV) tests, full step with the bf16 BEV image
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense_boxes.py tests/test_gpu_reference_dropin.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_f3_f4.log 2>&1; rc=$?; echo "== dense/boxes + reference drop-in exit $rc"; tail -3 gpurun_out/test_f3_f4.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_f3_f4.log | head -40; fi
timeout 600 python scripts/full_step.py > gpurun_out/full_step.json 2> gpurun_out/full_step.err; echo "full_step exit $?"; tail -3 gpurun_out/full_step.err | cut -c1-300; cat gpurun_out/full_step.json
