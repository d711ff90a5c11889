#!/bin/bash
# round 2, call A: spconv probe on the GPU box, the reference-through-dropins tests, the touched
# suites, smoke, bench.
mkdir -p gpurun_out
{
  echo "== spconv probe"; python -c "import spconv; print('spconv', spconv.__version__, spconv.__file__)" 2>&1 | tail -1
  python -c "import cumm; print('cumm', cumm.__file__)" 2>&1 | tail -1
  python -m pip list 2>/dev/null | grep -i -E "spconv|cumm" || echo "pip list: no spconv / cumm distribution"
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "spconv|cumm" || echo "wheelhouse: no spconv / cumm wheel"
  timeout 60 python -m pip download --no-deps -d /tmp/spw spconv-cu120 2>&1 | tail -2
  find / -xdev \( -iname "*spconv*" -o -iname "cumm*" \) -not -path "*/proc/*" -not -path "*com_b200*" -not -path "*/gpurun_out/*" 2>/dev/null | grep -v "$GRAFT_REPO_ROOT" | head -5
  echo "== probe done"
} > gpurun_out/spconv_probe.txt 2>&1
cat gpurun_out/spconv_probe.txt
for f in reference_dropin dense_boxes backbone; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_$f.log 2>&1; echo "== $f exit $?"; tail -5 gpurun_out/test_$f.log
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['breakdown_ms_per_step'].items()})
print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['layers'].items()})
PY
