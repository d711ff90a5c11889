#!/bin/bash
# round 2, call P: leaner conv_tr loops: parity, skeleton / fixed cost, real-rulebook ablation
mkdir -p gpurun_out
COMB_CONV_IMPL=tr timeout 400 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv_tr.log 2>&1; rc=$?; echo "== spconv (tr) exit $rc"; tail -3 gpurun_out/test_spconv_tr.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_spconv_tr.log | head -30; exit 1; fi
COMB_CONV_IMPL=tr timeout 120 python scripts/conv_fixed.py > gpurun_out/conv_fixed_tr.txt 2>&1; echo rc $?; cat gpurun_out/conv_fixed_tr.txt
COMB_CONV_IMPL=tr timeout 150 python scripts/conv_ablate.py > gpurun_out/conv_ablate_g_tr.txt 2>&1; echo rc $?; cat gpurun_out/conv_ablate_g_tr.txt
