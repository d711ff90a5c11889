#!/bin/bash
# ncu --set full captures (after 3 warm-up steps + 1 analysis step). Keeps gpurun_out small:
# conv: one launch each of <16,16>, <64,64>, <128,128>; memory-bound kernels: one step, exported as CSV.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
run() { timeout 900 $NCU "$@" python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_last.log 2>&1; echo "exit $?"; }
run -k regex:"spconv_tc_kernel<16, 16>" -s 20 -c 1 -o gpurun_out/prof_conv16 -f
run -k regex:"spconv_tc_kernel<64, 64>" -s 16 -c 1 -o gpurun_out/prof_conv64 -f
run -k regex:"spconv_tc_kernel<128, 128>" -s 20 -c 1 -o gpurun_out/prof_conv128 -f
run -k regex:"outset|nbrmap|vox_|dense_|hash_insert" -s 140 -c 35 -o /tmp/prof_mem -f
ncu -i /tmp/prof_mem.ncu-rep --page raw --csv > gpurun_out/prof_mem_raw.csv 2>/dev/null
for k in outset_emit outset_mark nbrmap vox_insert vox_mean dense_write; do
  ncu -i /tmp/prof_mem.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > gpurun_out/prof_mem_src_$k.csv 2>/dev/null
done
ls -la gpurun_out/ /tmp/prof_mem.ncu-rep; du -sh gpurun_out
