#!/bin/bash
# fifth (generic) class queue: whole GPU suite, timings of a scene that mixes specialised and generic BSDFs (sorted vs unsorted), bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_env5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_env5_pytest.log
tail -25 gpurun_out/r2_env5_pytest.log
for cfg in "f3mix 64 1024 0" "f3mix 64 1024 2" "f3mix 64 1024 256" "f3mix 64 1024 258" "envmap 64 1024 0"; do
  B2_RFILTER=gaussian python scripts/render_once.py $cfg 2>&1 | tail -1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_env5_bench.json 2> gpurun_out/r2_env5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_env5_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_env5_bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("value", d['value'], "e2e", d['e2e']['value'], "parity", {k:v['rel_l2'] for k,v in d.get('parity',{}).items()}, "envmap", d.get('envmap',{}).get('value'), d.get('envmap',{}).get('parity',{}).get('rel_l2'), "textured", d.get('textured',{}).get('value'), "trav", d.get('traversal',{}).get('mrays_s'), d.get('traversal',{}).get('commit_s'), {k:v.get('value') for k,v in d.get('configs',{}).items()})
PY
