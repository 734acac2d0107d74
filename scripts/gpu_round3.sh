#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for pool in 1048576 2097152 4194304 8388608; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool $pool > gpurun_out/bench_pool$pool.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pool$pool.json"))
r=d["roofline"]
print("pool=$pool value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", {k:round(v,1) for k,v in r["kernel_ms"].items()}, "ms/step", round(d["ms_per_step"],1), "dom", r["kernel"], r["avg_launch_ms"], r["avg_launch_ms_cuda_events"])
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity 1 > gpurun_out/bench_parity.json 2>> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_parity.json')); print('parity build value', d['value'], d['roofline']['kernel_ms'])"
tail -3 gpurun_out/bench.err
