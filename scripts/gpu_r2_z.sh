#!/bin/bash
# round 2, call Z: the wide-load conv_ts gather as the default: smoke, the conv parity file, the bench line of record
mkdir -p gpurun_out
timeout -s KILL 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r2.log
timeout -s KILL 200 python -m pytest tests/test_gpu_spconv.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "fwd_bf16 or single_tile or many_tiles_persistent or alternate" > gpurun_out/test_wide_default.log 2>&1; echo "spconv tests exit $?"; tail -2 gpurun_out/test_wide_default.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
print('conv TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3), d['roofline']['kernel'][:80])
PY
