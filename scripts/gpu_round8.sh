#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2>> gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cur.json"))
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", {k:round(v,1) for k,v in r["kernel_ms"].items()}, "ms/step", round(d["ms_per_step"],1), r["kernel"], round(r["achieved"],1), round(r["frac"],3))
PY
for pool in 2097152 8388608; do
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool $pool 2>> gpurun_out/bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pool $pool value', round(d['value'],1), d['roofline']['kernel_ms'])"
done
tail -3 gpurun_out/bench.err
