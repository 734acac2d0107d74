#!/bin/bash
# envmap, second pass: probes (fixed), scene-file + plugin routes, bench line with the envmap side measurement
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_envmap.py tests/test_gpu_xml.py tests/test_gpu_shim.py -q -p no:cacheprovider > gpurun_out/r2_env2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_env2_tests.log
tail -60 gpurun_out/r2_env2_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-traversal > gpurun_out/r2_env2_bench.json 2> gpurun_out/r2_env2_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2_env2_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_env2_bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("value", d['value'], "envmap", d.get('envmap'), "textured", d.get('textured',{}).get('value'))
PY
