#!/bin/bash
mkdir -p gpurun_out
true

timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu7_pytest.log
tail -4 gpurun_out/r2_gpu7_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-traversal --no-cpu-baseline > gpurun_out/r2_gpu7_bench.json 2> gpurun_out/r2_gpu7_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_gpu7_bench.json'))
print(d['value'], {k:v.get('rel_l2') for k,v in (d.get('parity') or {}).items()}, {k:(v.get('value'),v.get('ms_per_step')) for k,v in d['configs'].items()})
PY
