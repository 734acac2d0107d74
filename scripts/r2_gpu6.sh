#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_gpu6_bench.json 2> gpurun_out/r2_gpu6_bench.err
tail -3 gpurun_out/r2_gpu6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_gpu6_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('parity'), indent=1))
print(json.dumps(d.get('configs'), indent=1))
print(json.dumps(d.get('traversal'), indent=1))
print(json.dumps(d.get('cpu_baseline'), indent=1))
PY
