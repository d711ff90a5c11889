#!/bin/bash
# round 2, call S: f1 (target assignment + COM loss re-weighting on the device): parity, full config-2 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_center_targets.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_center_targets.log 2>&1; rc=$?; echo "== center targets exit $rc"; tail -3 gpurun_out/test_center_targets.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_center_targets.log | head -40; fi
timeout 600 python scripts/full_step.py > gpurun_out/full_step.json 2> gpurun_out/full_step.err; echo "full_step exit $?"; tail -3 gpurun_out/full_step.err; cat gpurun_out/full_step.json
