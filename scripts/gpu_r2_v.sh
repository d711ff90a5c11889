#!/bin/bash
# round 2, call V: velocity-head variant of the fused CenterHead post-processing (comb_centerhead_decode_nms_vel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_center_decode.py tests/test_gpu_reference_dropin.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/test_vel.log 2>&1; rc=$?; echo "== decode+dropin exit $rc"; tail -3 gpurun_out/test_vel.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_vel.log | head -30; fi
