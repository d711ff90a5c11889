#!/bin/bash
# round 2, last call: ncu --set full of the wide-load conv_ts at 64x64 (level 3 of the bench frames)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout -s KILL 150 $NCU -k regex:spconv_ts -s 3 -c 1 -f -o gpurun_out/prof_r2_conv64_wide python scripts/conv_one.py 3 64 64 > gpurun_out/ncu_conv64_wide.log 2>&1; echo "ncu conv 64x64 wide exit $?"; tail -2 gpurun_out/ncu_conv64_wide.log
