#!/bin/bash
# envmap on the device: its tests first, then the whole GPU suite, then a short headline bench (hot path must be unchanged)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_envmap.py -q -p no:cacheprovider > gpurun_out/r2_env1_envtests.log 2>&1
echo "env pytest rc=$?" >> gpurun_out/r2_env1_envtests.log
tail -40 gpurun_out/r2_env1_envtests.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_envmap.py > gpurun_out/r2_env1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_env1_pytest.log
tail -6 gpurun_out/r2_env1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-parity > gpurun_out/r2_env1_bench.json 2> gpurun_out/r2_env1_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_env1_bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("value", d['value'], "e2e", d['e2e']['value'], d['roofline']['kernel_ms'])
PY
