#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for fl in 64 0; do
  B2_FLAT_LIMIT=$fl python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat$fl.json 2> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_flat$fl.json"))
print("FLAT_LIMIT=$fl value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], "ms/step", d["ms_per_step"])
PY
done
for pool in 262144 524288 2097152; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool $pool > gpurun_out/bench_pool$pool.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pool$pool.json"))
print("pool=$pool value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "ms/step", d["ms_per_step"])
PY
done
tail -3 gpurun_out/bench.err
