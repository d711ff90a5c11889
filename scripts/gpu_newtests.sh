#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_dense_boxes.py tests/test_gpu_voxelize.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -k "train_mode or config4 or config5 or api_maximum" -s > gpurun_out/test_new.log 2>&1; echo "== new tests exit $?"; tail -15 gpurun_out/test_new.log
