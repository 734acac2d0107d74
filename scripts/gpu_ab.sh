#!/bin/bash
mkdir -p gpurun_out
for v in "" _mb5 _mb7 _nofm; do
  export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts$v.so
  echo "=== variant '$v'"
  python -m pytest tests -m gpu -q -x 2>&1 | tail -2
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab$v.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab$v.json"))
r=d["roofline"]
print("value", round(d["value"],1), "kernel_ms", {k:round(v,1) for k,v in r["kernel_ms"].items()}, "ms/step", round(d["ms_per_step"],1))
PY
done
