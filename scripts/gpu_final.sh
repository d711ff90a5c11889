#!/bin/bash
# round-end evidence: full GPU parity suite + smoke, bench (both arms), ncu launch list, ncu --set full of the dominant kernel
bash scripts/gpu_full.sh
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
for w in 16 64 128; do
  timeout 600 $NCU -k regex:"spconv_ts_kernel.*\)$w, .*\)$w," -s 12 -c 1 -o gpurun_out/prof_ts$w -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ts$w.log 2>&1; echo "ncu ts$w exit $?"
done
timeout 600 $NCU -k regex:"nbrmap_indexed|vox_insert|vox_mean|index_emit|index_count|dense_scatter" -s 60 -c 24 -o gpurun_out/prof_mem -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_mem.log 2>&1; echo "ncu mem exit $?"
ls -la gpurun_out/*.ncu-rep
