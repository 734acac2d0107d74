#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu8_pytest.log
tail -4 gpurun_out/r2_gpu8_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_gpu8_bench.json 2> gpurun_out/r2_gpu8_bench.err
tail -3 gpurun_out/r2_gpu8_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_gpu8_bench.json'))
print(d['value'], {k:v.get('rel_l2') for k,v in (d.get('parity') or {}).items()}, {k:(v.get('value'),v.get('ms_per_step')) for k,v in d['configs'].items()})
print(d.get('traversal'))
PY
B2_NO_WIDE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-configs > gpurun_out/r2_gpu8_bench_bvh2.json 2> gpurun_out/r2_gpu8_bench_bvh2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_gpu8_bench_bvh2.json')); print('BVH2', d.get('traversal'))"
