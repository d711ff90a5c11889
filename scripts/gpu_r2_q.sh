#!/bin/bash
# round 2, call Q: one polling lane per waiting warp (both tcgen05 conv kernels): parity, ablation, bench
mkdir -p gpurun_out
for impl in tr ts; do
COMB_CONV_IMPL=$impl timeout 400 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "fwd_bf16 or persistent" > gpurun_out/test_spconv_$impl.log 2>&1; rc=$?; echo "== spconv ($impl) exit $rc"; tail -2 gpurun_out/test_spconv_$impl.log
if [ $rc != 0 ]; then grep -E "^E |Error|Timeout|assert" gpurun_out/test_spconv_$impl.log | head -30; exit 1; fi
COMB_CONV_IMPL=$impl timeout 150 python scripts/conv_ablate.py > gpurun_out/conv_ablate_g_$impl.txt 2>&1; echo rc $?; cat gpurun_out/conv_ablate_g_$impl.txt
done
