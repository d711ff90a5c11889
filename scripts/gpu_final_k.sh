#!/bin/bash
# round-end evidence (r1_k): full GPU parity suite + smoke, bench (both arms), ncu launch list, ncu --set full of the
# tcgen05 wgrad kernel (one launch of the 16-, 64- and 128-channel variants inside the bench's train leg)
bash scripts/gpu_full.sh
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:"spconv_wgrad_tc_kernel" -s 42 -c 21 -o gpurun_out/prof_wgrad -f python scripts/train_profile.py bf16 noprof > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad exit $?"
ls -la gpurun_out/*.ncu-rep
